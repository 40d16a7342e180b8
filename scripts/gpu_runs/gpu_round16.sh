#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python scripts/gpu_exp_tile.py mid > gpurun_out/exp_tile_mid4.log 2>&1; grep tile gpurun_out/exp_tile_mid4.log
timeout 900 python scripts/gpu_exp_tile.py c4 > gpurun_out/exp_tile_c4_4.log 2>&1; grep tile gpurun_out/exp_tile_c4_4.log
