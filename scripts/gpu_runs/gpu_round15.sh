#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python scripts/gpu_exp_1d_nondefault.py > gpurun_out/exp_1d_nondefault.log 2>&1; cat gpurun_out/exp_1d_nondefault.log
