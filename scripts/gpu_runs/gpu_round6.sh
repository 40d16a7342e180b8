#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --workload mid --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_mid_pol.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/bench_mid_pol.json')); print('mid', d['value'])"
timeout 1200 python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_pol.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/bench_c4_pol.json')); print('c4', d['value'])"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:pgsgd_iter -s 1 -c 1 -o gpurun_out/prof_c4_pol -f python bench.py --workload c4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c4_pol.log 2>&1
