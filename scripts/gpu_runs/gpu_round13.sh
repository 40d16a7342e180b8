#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 600 python scripts/gpu_exp_1d.py mid > gpurun_out/exp_1d_mid.log 2>&1; cat gpurun_out/exp_1d_mid.log
timeout 600 python scripts/gpu_exp_1d.py c4 > gpurun_out/exp_1d_c4.log 2>&1; cat gpurun_out/exp_1d_c4.log
