#!/bin/bash
# Round 2, call 18 (1 GPU): the end-to-end chain at c4 scale on one GFA file (3.6 GB): `pgsgd layout --device-ingest` (file -> .lay),
# odgi + our shim (odgi ingest, graph_t walk, upload, run), and the reference's own CUDA path live on the same file.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r02_c18_prep.log 2>&1
import sys, time
sys.path.insert(0, '.')
from odgi_b200 import synth, graphio
t = time.time(); g = synth.preset("c4"); print("gen", round(time.time() - t, 1), g.N, g.S, flush=True)
t = time.time(); graphio.save_graph_arrays("/tmp/c4.arr", g); print("arr", round(time.time() - t, 1), flush=True)
PY
T0=$(date +%s.%N); scripts/probes/arr2gfa /tmp/c4.arr /tmp/c4.gfa >> gpurun_out/r02_c18_prep.log 2>&1; T1=$(date +%s.%N); python -c "print('arr2gfa wall', round($T1 - $T0, 1), 's')" >> gpurun_out/r02_c18_prep.log
rm -f /tmp/c4.arr; ls -l /tmp/c4.gfa >> gpurun_out/r02_c18_prep.log; nproc >> gpurun_out/r02_c18_prep.log
{
odgi_b200/host/pgsgd layout -i /tmp/c4.gfa -o /tmp/c4_dev.lay --gpu --init-seed 42 --device-ingest --timing 2>&1 | grep -E '^\{'
odgi_b200/host/pgsgd layout -i /tmp/c4.gfa -o /tmp/c4_dev.lay --gpu --init-seed 42 --device-ingest --timing 2>&1 | grep -E '^\{'
ls -l /tmp/c4_dev.lay | awk '{print "{\"file\": \"" $9 "\", \"bytes\": " $5 "}"}'
(cd /tmp && PGSGD_SHIM_TIMING=1 timeout 900 "$GRAFT_REPO_ROOT"/oracle/_ref/shim_driver layout /tmp/c4.gfa - 30 32 2>&1 | grep -E '^\{')
(cd /tmp && timeout 900 "$GRAFT_REPO_ROOT"/oracle/_ref/ref_gpu_driver /tmp/c4.gfa - 30 32 2>/dev/null | grep -E '^\{')
} > gpurun_out/r02_c18_e2e_c4.jsonl 2>&1
cat gpurun_out/r02_c18_prep.log gpurun_out/r02_c18_e2e_c4.jsonl
