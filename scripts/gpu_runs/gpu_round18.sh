#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
# experiment: 5 CTAs/SM build of the tile kernel
cp odgi_b200/libpgsgd_b200.so /tmp/keep.so; cp odgi_b200/libpgsgd_b5_experiment.so odgi_b200/libpgsgd_b200.so
timeout 600 python scripts/gpu_exp_tile.py c4 2>&1 | grep -E "tile    batch=1"
cp /tmp/keep.so odgi_b200/libpgsgd_b200.so
