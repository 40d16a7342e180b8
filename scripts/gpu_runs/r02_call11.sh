#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python scripts/gpu_exp_1d_streams.py mid > gpurun_out/r02_c11_exp_1d_streams_mid.log 2>&1
cat gpurun_out/r02_c11_exp_1d_streams_mid.log
