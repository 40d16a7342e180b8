#!/bin/bash
mkdir -p gpurun_out
for gran in 32 64 128; do
  PGSGD_L2_FETCH=$gran timeout 600 python scripts/gpu_exp_occupancy.py mid > gpurun_out/exp_occ_mid_l2f$gran.log 2>&1
done
PGSGD_L2_FETCH=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pgsgd_iter -s 1 -c 1 -o gpurun_out/prof_mid_b4_l2f32 -f python bench.py --workload mid --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_mid_b4_l2f32.log 2>&1
tail -n 40 gpurun_out/exp_occ_mid_l2f32.log
