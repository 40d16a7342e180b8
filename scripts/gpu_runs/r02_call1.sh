#!/bin/bash
# Round 2, call 1 (1 GPU): the GPU suite on the new default (pipelined) tile kernel, the kernel-variant sweep on c4 with the
# quality readout, and ncu --set full captures of the new kernel (one early iteration, one cooling iteration).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_runs/r02_call1.sh'
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_c1_gpu.txt 2>&1
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r02_c1_pytest.log
timeout 600 python scripts/gpu_exp_tile2.py c4 --quality > gpurun_out/r02_c1_exp_tile2_c4.log 2>&1
timeout 300 python scripts/gpu_exp_tile2.py mid > gpurun_out/r02_c1_exp_tile2_mid.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pgsgd_tile2 -s 1 -c 1 -f -o gpurun_out/r02_prof_tile2_c4_early \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_c1_ncu_early.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pgsgd_tile2 -s 21 -c 1 -f -o gpurun_out/r02_prof_tile2_c4_cool \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_c1_ncu_cool.log 2>&1
tail -15 gpurun_out/r02_c1_pytest.log; cat gpurun_out/r02_c1_exp_tile2_c4.log gpurun_out/r02_c1_exp_tile2_mid.log; tail -3 gpurun_out/r02_c1_ncu_early.log
