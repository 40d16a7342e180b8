"""GPU experiment (round 2): where does the summed coordinate write (red.add) really turn unstable on ONE GPU?
LPA 1D (max node depth 244, S = 202 806) with PGSGD_FLAG_KEEP_ADD (hub safeguard off) and the exchange write, over the number of
worker streams: concurrent terms per hub-node end = streams * 2 * 244 / S.  The oracle's worst-case model
(scripts/cpu_exp_inflight_model.py) puts the onset at ~2.5; measured so far: in band at 2.3, diverged at 9."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import odgi_b200  # noqa: E402
from odgi_b200 import capi  # noqa: E402
from odgi_b200.arrays import read_arrays  # noqa: E402
from oracle import oracle as orc  # noqa: E402

a = read_arrays(os.path.join(ROOT, "tests", "golden", "LPA.graph.arr.gz"))
gd, go = odgi_b200.graph_from_arrays(a), orc.Graph.from_arrays(a)
band = json.load(open(os.path.join(ROOT, "tests", "golden", "stress_reference.json")))["LPA.sort1d"]
print(f"LPA 1D reference band {band['mean']:.4f} +- {band['sd']:.4f}", flush=True)
with odgi_b200.Engine(gd) as e:
    for name, flags in (("red.add", capi.PGSGD_FLAG_KEEP_ADD), ("exch", capi.PGSGD_FLAG_EXCH_WRITE)):
        for streams in (512, 768, 1024, 1280, 1536, 1792, 2048, 3072, 4096):
            cd = capi.sort_defaults(gd, n_streams=streams, batch=1, flags=flags, sampling=capi.SAMPLING_STREAM)
            e.set_coords_1d(None)
            e.run_1d(cd)
            x = e.get_coords_1d()
            s = orc.path_stress_1d(go, x, band["n_pairs"], band["seed"]) if np.all(np.isfinite(x)) else float("nan")
            print(f"  {name:8s} streams {streams:5d}  terms per hub end {streams * 2 * 244 / go.S:5.2f}  stress {s:.4f}  ({100 * (s / band['mean'] - 1):+.1f} %)", flush=True)
