#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/gpu_sweep_streams.py > gpurun_out/sweep_streams.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pgsgd_iter -s 1 -c 2 -o gpurun_out/prof_mid_b4 -f python bench.py --workload mid --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_mid_b4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pgsgd_iter -s 1 -c 1 -o gpurun_out/prof_mid_b1 -f python bench.py --workload mid --steps 2 --warmup 1 --batch 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_mid_b1.log 2>&1
timeout 1200 python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
tail -5 gpurun_out/sweep_streams.log; cut -c1-300 gpurun_out/bench_c4.json
