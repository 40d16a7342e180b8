"""GPU diagnostic (round 2): 1D on `mid` — every device variant (reference-exact stream sampling included, which is bit-identical
to the oracle for ONE stream) ends ~9 % above the reference's far-stress band (6 CPU threads).  Is it the number of concurrent
worker streams (Hogwild staleness)?  Final far / local stress vs n_streams."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "mid"
g = synth.preset(wl)
band = json.load(open(os.path.join(ROOT, "tests", "golden", "stress_reference_scale.json")))[f"{wl}.sort1d"]
print(f"workload={wl} N={g.N} S={g.S}  reference band far {band['far']['mean']:.6g} +- {band['far']['sd']:.3g}  local {band['local']['mean']:.5g} +- {band['local']['sd']:.3g}", flush=True)
with odgi_b200.Engine(g) as e:
    counts = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 65536, 16384, 4096, 1024, 256]
    seeds = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [9399220, 7]
    for n_streams in counts:
        far, loc, rate = [], [], []
        for seed in seeds:
            cd = capi.sort_defaults(g, seed=seed, sampling=1, n_streams=n_streams)
            e.set_coords_1d(None)
            st = e.run_1d(cd)
            far.append(e.path_stress(1, 4_000_000, 12345)); loc.append(e.local_stress(1, 4_000_000, 12345))
            rate.append(st["term_updates"] / st["seconds_iterations"] / 1e9)
        print(f"stream sampling, n_streams {n_streams or 'default':>8}: far {np.mean(far):.6g} sd {np.std(far, ddof=1):.2g} [{' '.join(f'{v:.4g}' for v in far)}]  local {np.mean(loc):.5g}   {np.mean(rate):.2f} G/s", flush=True)
