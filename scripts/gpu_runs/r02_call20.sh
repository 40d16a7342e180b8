#!/bin/bash
# Round 2, call 20 (1 GPU): window order of the tile visits (scripts/gpu_exp_window.py) on c4, c4x and mid.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
timeout 600 python scripts/gpu_exp_window.py c4
timeout 600 python scripts/gpu_exp_window.py c4x
timeout 600 python scripts/gpu_exp_window.py mid 9399220,1234567,42
} > gpurun_out/r02_c20_window_order.log 2>&1
cat gpurun_out/r02_c20_window_order.log
