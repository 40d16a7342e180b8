#!/bin/bash
# diagnostic: run-to-run spread of the single-process 2-GPU layout through the reference call chain (shim_driver) and through ctypes
cd "$GRAFT_REPO_ROOT"
python - <<'PY' > gpurun_out/multi_inproc_diag.log 2>&1
import os, subprocess, sys, json, tempfile
import numpy as np
sys.path.insert(0, os.getcwd())
import odgi_b200
from odgi_b200 import capi, synth
from odgi_b200.arrays import read_arrays
from oracle import oracle as orc
a = read_arrays("tests/golden/DRB1-3123.graph.arr.gz")
gd, go = odgi_b200.graph_from_arrays(a), orc.Graph.from_arrays(a)
ROOT0 = os.getcwd()
tmp = tempfile.mkdtemp()
gfa = os.path.join(tmp, "g.gfa"); synth.write_gfa(gd, gfa)
def shim(env_extra, tag):
    out = os.path.join(tmp, "o.arr")
    subprocess.run([os.path.join(ROOT0, "oracle/_ref/shim_driver"), "layout", gfa, out], check=True, cwd=tmp, env=dict(os.environ, **env_extra), stdout=subprocess.DEVNULL)
    r = read_arrays(out)
    print(tag, env_extra, "stress", orc.path_stress_2d(go, r["X"], r["Y"], 1000000, 12345), "init", orc.path_stress_2d(go, r["X0"], r["Y0"], 1000000, 12345), flush=True)
    return r
for i in range(2): shim({}, "shim-1gpu")
for i in range(4): shim({"PGSGD_GPUS": "2"}, "shim-hybrid")
for i in range(2): shim({"PGSGD_GPUS": "2", "PGSGD_MULTI": "peer"}, "shim-peer")
for i in range(2): shim({"PGSGD_GPUS": "2", "PGSGD_MULTI": "allreduce"}, "shim-allreduce")
r = shim({}, "shim-1gpu")
X0, Y0 = r["X0"], r["Y0"]          # the reference's initialisation, replayed through ctypes
for mode, name in ((capi.MULTI_HYBRID, "hybrid"), (capi.MULTI_PEER, "peer"), (capi.MULTI_ALLREDUCE, "allreduce")):
    for i in range(3):
        X, Y, st = odgi_b200.layout_2d_multi(gd, capi.layout_defaults(gd), X0, Y0, 2, mode)
        print("ctypes", name, "shim-init", orc.path_stress_2d(go, X, Y, 1000000, 12345), st["term_updates"], flush=True)
X0, Y0 = orc.layout_init(go, 42)
for i in range(3):
    X, Y, st = odgi_b200.layout_2d_multi(gd, capi.layout_defaults(gd), X0, Y0, 2, capi.MULTI_HYBRID)
    print("ctypes hybrid seed42-init", orc.path_stress_2d(go, X, Y, 1000000, 12345), flush=True)
PY
tail -40 gpurun_out/multi_inproc_diag.log
