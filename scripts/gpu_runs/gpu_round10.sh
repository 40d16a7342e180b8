#!/bin/bash
mkdir -p gpurun_out
make -C odgi_b200/host > /dev/null 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-250 gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/bench_reference.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_c4.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:pgsgd_tile -s 1 -c 1 -o gpurun_out/prof_c4_tile2 -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c4_tile2.log 2>&1
timeout 900 python bench.py --workload chr6.C4 --steps 27 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chr6C4.json 2>/dev/null; cut -c1-200 gpurun_out/bench_chr6C4.json
