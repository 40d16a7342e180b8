#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python scripts/gpu_exp_longthin.py longthin > gpurun_out/r02_c4_exp_longthin.log 2>&1
timeout 600 python scripts/gpu_exp_longthin.py mid 9399220,1234567,42 > gpurun_out/r02_c4_exp_mid.log 2>&1
cat gpurun_out/r02_c4_exp_longthin.log gpurun_out/r02_c4_exp_mid.log
