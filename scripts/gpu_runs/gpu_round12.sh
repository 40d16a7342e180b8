#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python scripts/gpu_exp_tile.py mid > gpurun_out/exp_tile_mid3.log 2>&1; cat gpurun_out/exp_tile_mid3.log
