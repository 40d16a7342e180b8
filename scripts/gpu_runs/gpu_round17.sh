#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gpu_exp_tile.py mid > gpurun_out/exp_tile_mid5.log 2>&1; grep -E "tile|stream" gpurun_out/exp_tile_mid5.log
timeout 900 python scripts/gpu_exp_tile.py c4 > gpurun_out/exp_tile_c4_5.log 2>&1; grep -E "tile|stream" gpurun_out/exp_tile_c4_5.log
