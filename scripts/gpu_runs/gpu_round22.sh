#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python scripts/gpu_exp_tile.py mid 2>&1 | grep -E "tile    batch=1|stream"
timeout 600 python scripts/gpu_exp_tile_small.py 2>&1 | grep -E "auto"
