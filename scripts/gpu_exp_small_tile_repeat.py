"""GPU check (round 2): run-to-run spread of the small-graph tile-sampling parity test (tests/test_gpu_parity.py
test_2d_tile_sampling_stress_within_reference_band): the same three seeds, repeated — how close to the 1 % gate do single runs come?"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi  # noqa: E402
from odgi_b200.arrays import read_arrays  # noqa: E402
from oracle import oracle as orc  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bands = json.load(open(os.path.join(ROOT, "tests", "golden", "stress_reference.json")))
for name in ("DRB1-3123", "chr6.C4"):
    a = read_arrays(os.path.join(ROOT, "tests", "golden", f"{name}.graph.arr.gz"))
    gd, go = odgi_b200.graph_from_arrays(a), orc.Graph.from_arrays(a)
    band = bands[f"{name}.layout2d"]
    X0, Y0 = orc.layout_init(go, seed=42)
    for sampling in (capi.SAMPLING_TILE, capi.SAMPLING_STREAM):
        allv = []
        for rep in range(reps):
            vals = []
            for seed in (9399220, 1234567, 42):
                X, Y, st = odgi_b200.layout_2d(gd, capi.layout_defaults(gd, seed=seed, sampling=sampling), X0, Y0)
                vals.append(orc.path_stress_2d(go, X, Y, n_pairs=band["n_pairs"], seed=band["seed"]))
            allv.append(vals)
        v = np.array(allv)
        print(f"{name} sampling={sampling} band {band['mean']:.5f} +- {band['sd']:.5f}  per seed mean {v.mean(0).round(5)} min {v.min(0).round(5)} max {v.max(0).round(5)}  "
              f"worst relative deviation {100 * np.max(np.abs(v / band['mean'] - 1)):.2f} %", flush=True)
