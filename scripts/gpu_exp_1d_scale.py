"""GPU diagnostic (round 2): 1D (`odgi sort -Y`) at scale — pipelined tile kernel vs legacy tile kernel vs stream sampling on
`mid`, final far / local stress over seeds, against the reference band (tests/golden/stress_reference_scale.json)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "mid"
g = synth.preset(wl) if wl in synth.PRESETS else synth.generate(3_000_000, 6, seed=42)
band = json.load(open(os.path.join(ROOT, "tests", "golden", "stress_reference_scale.json"))).get(f"{wl}.sort1d")
print(f"workload={wl} N={g.N} S={g.S}  reference band far {band['far']['mean']:.6g} +- {band['far']['sd']:.3g}  local {band['local']['mean']:.5g} +- {band['local']['sd']:.3g}" if band else wl, flush=True)
with odgi_b200.Engine(g) as e:
    for name, kw in (("tile2", dict(sampling=2)), ("legacy tile", dict(sampling=2, flags=capi.FLAG_LEGACY_TILE)), ("stream", dict(sampling=1)),
                     ("tile2 exch", dict(sampling=2, flags=capi.FLAG_EXCH_WRITE))):
        far, loc, rate = [], [], []
        for seed in (9399220, 1234567, 42, 7):
            cd = capi.sort_defaults(g, seed=seed, **kw)
            e.set_coords_1d(None)
            st = e.run_1d(cd)
            far.append(e.path_stress(1, 4_000_000, 12345)); loc.append(e.local_stress(1, 4_000_000, 12345))
            rate.append(st["term_updates"] / st["seconds_iterations"] / 1e9)
        print(f"{name:12s} far mean {np.mean(far):.6g} sd {np.std(far, ddof=1):.3g} [{' '.join(f'{v:.4g}' for v in far)}]  local mean {np.mean(loc):.5g} sd {np.std(loc, ddof=1):.3g}"
              f"   {np.mean(rate):.1f} G/s  flags_used {st['flags_used']}", flush=True)
