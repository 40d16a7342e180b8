#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python scripts/gpu_exp_tile.py mid > gpurun_out/exp_tile_mid.log 2>&1; cat gpurun_out/exp_tile_mid.log
timeout 900 python scripts/gpu_exp_tile.py c4 > gpurun_out/exp_tile_c4.log 2>&1; cat gpurun_out/exp_tile_c4.log
