#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python scripts/gpu_exp_1d_scale.py mid > gpurun_out/r02_c10_exp_1d_mid.log 2>&1
timeout 600 python scripts/gpu_exp_sweeps.py c4 > gpurun_out/r02_c10_exp_sweeps_c4.log 2>&1
timeout 400 python scripts/gpu_exp_sweeps.py mid > gpurun_out/r02_c10_exp_sweeps_mid.log 2>&1
cat gpurun_out/r02_c10_exp_1d_mid.log gpurun_out/r02_c10_exp_sweeps_c4.log gpurun_out/r02_c10_exp_sweeps_mid.log
