#!/bin/bash
# first GPU contact: parity tests, smoke, and a first bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
for b in 1 2 4; do
  timeout 600 python bench.py --workload mid --steps 6 --warmup 3 --batch $b --no-e2e --no-cpu-baseline > gpurun_out/bench_mid_b$b.json 2> gpurun_out/bench_mid_b$b.err
done
timeout 900 python bench.py --workload mid --steps 6 --warmup 3 > gpurun_out/bench_mid_full.json 2> gpurun_out/bench_mid_full.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench_mid_b*.json | cut -c1-400
