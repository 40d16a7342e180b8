#!/bin/bash
mkdir -p gpurun_out
make -C odgi_b200/host > /dev/null 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['e2e'], d.get('quality'))"; tail -3 gpurun_out/bench_default.err
