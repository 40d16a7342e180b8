#!/bin/bash
# Round 2, call 21 (1 GPU): final validation of the library as committed — memcheck, the whole GPU suite (scale bands included),
# smoke, the official bench line, the reference arm, the ncu launch list of the bench command.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/gpu_sanitize.py > gpurun_out/r02_c21_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_c21_memcheck.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_c21_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c21_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r02_c21_bench_c4_n1.json 2> gpurun_out/r02_c21_bench_c4_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c21_bench_reference_arm.json 2> gpurun_out/r02_c21_bench_reference_arm.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c21_launches_c4.csv \
    python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_c21_launches.log 2>&1
python - <<'PY' > gpurun_out/r02_c21_prep.log 2>&1
import sys, time
sys.path.insert(0, '.')
from odgi_b200 import synth, graphio
g = synth.preset("c4"); graphio.save_graph_arrays("/tmp/c4.arr", g)
PY
scripts/probes/arr2gfa /tmp/c4.arr /tmp/c4.gfa >> gpurun_out/r02_c21_prep.log 2>&1; rm -f /tmp/c4.arr
for i in 1 2 3; do odgi_b200/host/pgsgd layout -i /tmp/c4.gfa -o /tmp/c4_dev.lay --gpu --init-seed 42 --device-ingest --timing 2>&1 | grep -E '^\{'; done > gpurun_out/r02_c21_ingest_c4.jsonl 2>&1
rm -f /tmp/c4.gfa /tmp/c4_dev.lay
cat gpurun_out/r02_c21_ingest_c4.jsonl
tail -4 gpurun_out/r02_c21_memcheck.log; tail -12 gpurun_out/r02_c21_pytest_gpu.log; tail -3 gpurun_out/r02_c21_smoke.log
head -c 2500 gpurun_out/r02_c21_bench_c4_n1.json; echo; head -c 1200 gpurun_out/r02_c21_bench_reference_arm.json
