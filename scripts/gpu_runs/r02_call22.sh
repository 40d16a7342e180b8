#!/bin/bash
# Round 2, call 22 (1 GPU): device ingest at c4 scale again, with the line scan on host threads and the staged text upload;
# the ingest tests.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gfa_ingest.py tests/test_gpu_host.py -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_c22_pytest.log
python - <<'PY' > gpurun_out/r02_c22_prep.log 2>&1
import sys, time
sys.path.insert(0, '.')
from odgi_b200 import synth, graphio
t = time.time(); g = synth.preset("c4"); print("gen", round(time.time() - t, 1), g.N, g.S, flush=True)
graphio.save_graph_arrays("/tmp/c4.arr", g)
PY
scripts/probes/arr2gfa /tmp/c4.arr /tmp/c4.gfa >> gpurun_out/r02_c22_prep.log 2>&1; rm -f /tmp/c4.arr
{
odgi_b200/host/pgsgd scan -i /tmp/c4.gfa
PGSGD_SCAN_THREADS=1 odgi_b200/host/pgsgd scan -i /tmp/c4.gfa
for i in 1 2 3; do odgi_b200/host/pgsgd layout -i /tmp/c4.gfa -o /tmp/c4_dev.lay --gpu --init-seed 42 --device-ingest --timing 2>&1 | grep -E '^\{'; done
PGSGD_UPLOAD_THREADS=1 odgi_b200/host/pgsgd layout -i /tmp/c4.gfa -o /tmp/c4_dev.lay --gpu --init-seed 42 --device-ingest --timing 2>&1 | grep -E '^\{'
PGSGD_UPLOAD_THREADS=12 odgi_b200/host/pgsgd layout -i /tmp/c4.gfa -o /tmp/c4_dev.lay --gpu --init-seed 42 --device-ingest --timing 2>&1 | grep -E '^\{'
} > gpurun_out/r02_c22_ingest_c4.jsonl 2>&1
cat gpurun_out/r02_c22_pytest.log gpurun_out/r02_c22_prep.log gpurun_out/r02_c22_ingest_c4.jsonl
