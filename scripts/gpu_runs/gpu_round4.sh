#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gpu_exp_store.py mid > gpurun_out/exp_store_mid.log 2>&1
for f in 0 1; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pgsgd_iter -s 1 -c 1 -o gpurun_out/prof_mid_b1_f$f -f python bench.py --workload mid --steps 2 --warmup 1 --batch 1 --streams 113664 --flags $f --no-e2e --no-cpu-baseline > gpurun_out/ncu_mid_b1_f$f.log 2>&1
done
cat gpurun_out/exp_store_mid.log
