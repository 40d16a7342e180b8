#!/bin/bash
# Round 2, call 8 (1 GPU): the new GPU tests, smoke(), the reference call chain end to end (shim phase timing, flatten thread
# scaling) next to the device ingest and to the reference's own CUDA path on the same GFA, the official bench line, the ncu
# launch list of the bench command.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_c8_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c8_smoke.log 2>&1
python - <<'PY' > gpurun_out/r02_c8_prep.log 2>&1
import sys, time
sys.path.insert(0, '.')
from odgi_b200 import synth
t = time.time(); g = synth.preset("mid"); print("gen", time.time() - t, g.N, g.S)
t = time.time(); synth.write_gfa(g, "/tmp/mid.gfa"); print("gfa", time.time() - t)
PY
{
for T in 1 8 32; do
  (cd /tmp && PGSGD_SHIM_TIMING=1 timeout 600 "$GRAFT_REPO_ROOT"/oracle/_ref/shim_driver layout /tmp/mid.gfa - 30 $T 2>&1 | grep -E '^\{')
done
( /usr/bin/time -f '{"cli": "pgsgd layout (host GFA parse)", "wall_s": %e}' odgi_b200/host/pgsgd layout -i /tmp/mid.gfa -o /tmp/mid_host.lay --gpu --init-seed 42 ) 2>&1 | grep -E '^\{'
odgi_b200/host/pgsgd layout -i /tmp/mid.gfa -o /tmp/mid_dev.lay --gpu --init-seed 42 --device-ingest --timing 2>&1 | grep -E '^\{'
ls -l /tmp/mid_host.lay /tmp/mid_dev.lay | awk '{print "{\"file\": \"" $9 "\", \"bytes\": " $5 "}"}'
(cd /tmp && timeout 900 "$GRAFT_REPO_ROOT"/oracle/_ref/ref_gpu_driver /tmp/mid.gfa - 30 32 2>/dev/null | grep -E '^\{')
} > gpurun_out/r02_c8_shim_e2e_mid.jsonl 2>&1
timeout 600 python bench.py > gpurun_out/r02_c8_bench_c4_n1.json 2> gpurun_out/r02_c8_bench_c4_n1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c4.csv \
    python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_c8_launches.log 2>&1
tail -12 gpurun_out/r02_c8_pytest_gpu.log; cat gpurun_out/r02_c8_smoke.log | tail -3; cat gpurun_out/r02_c8_prep.log gpurun_out/r02_c8_shim_e2e_mid.jsonl; head -c 3000 gpurun_out/r02_c8_bench_c4_n1.json
