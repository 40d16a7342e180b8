#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python scripts/gpu_exp_longthin.py longthin 9399220,1234567,42,7,8,9,10,11,12,13 > gpurun_out/r02_c5_exp_longthin_feistel.log 2>&1
timeout 600 python scripts/gpu_exp_tile2.py c4 > gpurun_out/r02_c5_exp_tile2_c4.log 2>&1
timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_c5_pytest_all.log
cat gpurun_out/r02_c5_exp_longthin_feistel.log gpurun_out/r02_c5_exp_tile2_c4.log; tail -8 gpurun_out/r02_c5_pytest_all.log
