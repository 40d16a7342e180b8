"""GPU record (round 2): the far / local stress values of the default 1D runs tests/test_gpu_scale.py gates (same seeds), so that a
band update can be checked against them without a GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

for name, gen in (("longthin", (3_000_000, 6)), ("mid", (500_000, 90))):
    g = synth.generate(*gen, seed=42)
    far, loc = [], []
    with odgi_b200.Engine(g) as e:
        for seed in (9399220, 1234567, 42):
            e.set_coords_1d(None)
            e.run_1d(capi.sort_defaults(g, seed=seed))
            far.append(e.path_stress(1, 4_000_000, 12345)); loc.append(e.local_stress(1, 4_000_000, 12345))
    print(json.dumps({"graph": name, "kind": "sort1d", "seeds": [9399220, 1234567, 42], "far": far, "local": loc}), flush=True)
