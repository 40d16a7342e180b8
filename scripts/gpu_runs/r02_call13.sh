#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python scripts/gpu_exp_1d_streams.py mid 0,65536,32768,16384,8192 9399220,7,8,9,10,11,12,13,14,15 > gpurun_out/r02_c13_exp_1d_streams_mid.log 2>&1
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/r02_c13_pytest_gpu.log
cat gpurun_out/r02_c13_exp_1d_streams_mid.log; tail -6 gpurun_out/r02_c13_pytest_gpu.log
