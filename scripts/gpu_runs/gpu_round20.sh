#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['quality']['stress_final'], d['roofline']['frac'], d['cpu_baseline'])"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_c4.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:pgsgd_tile -s 1 -c 1 -o gpurun_out/prof_c4_tile3 -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c4_tile3.log 2>&1
grep -c pgsgd gpurun_out/launches_c4.csv
