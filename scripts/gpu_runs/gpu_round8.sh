#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:pgsgd_tile -s 1 -c 1 -o gpurun_out/prof_c4_tile -f python bench.py --workload c4 --steps 2 --warmup 1 --batch 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c4_tile.log 2>&1
tail -3 gpurun_out/ncu_c4_tile.log
