"""GPU diagnostic (round 2): why does the default tile-mode run on `longthin` (3.6e6 nodes, 6 haplotypes, 4e7 bp) end above the
CPU reference's far-stress band (0.000505 +- 0.000016, five reference runs) and with a large seed-to-seed spread?
Final far / local stress over several seeds for: the pipelined tile kernel (default), the legacy tile kernel, stream sampling
(the reference-exact sampler), fewer resident CTAs (less Hogwild staleness), sweep order, the exchange write."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "longthin"
seeds = [9399220, 1234567, 42, 7, 8, 9] if len(sys.argv) < 3 else [int(s) for s in sys.argv[2].split(",")]
g = synth.preset(wl) if wl in synth.PRESETS else synth.generate(3_000_000, 6, seed=42)
X0, Y0 = odgi_b200.layout_init(g, 42)
print(f"workload={wl} N={g.N} S={g.S}", flush=True)
VARIANTS = [("tile2 default", dict(sampling=2)), ("legacy tile", dict(sampling=2, flags=capi.FLAG_LEGACY_TILE)), ("stream sampling", dict(sampling=1)),
            ("tile2 148 CTAs", dict(sampling=2, n_streams=148 * 256)), ("tile2 37 CTAs", dict(sampling=2, n_streams=37 * 256)),
            ("tile2 sweep", dict(sampling=2, flags=capi.FLAG_SWEEP_TILES)), ("tile2 exch write", dict(sampling=2, flags=capi.FLAG_EXCH_WRITE)),
            ("stream 37 CTAs", dict(sampling=1, n_streams=37 * 256))]
if len(sys.argv) > 3 and sys.argv[3] == "sampling":   # the with-replacement / scrambled variants of the tile sampling (scripts/gpu_exp_mid2d_seeds.py)
    M = capi.FLAG_X_STEP_SCRAMBLE
    VARIANTS = [("tile2 default", dict(sampling=2)), ("stream sampling", dict(sampling=1)), ("tile2 scrambled lanes", dict(sampling=2, flags=M)),
                ("tile2 scrambled pairs", dict(sampling=2, flags=M | capi.FLAG_X_SCRAMBLE_PAIRS)),
                ("tile2 tiles+steps w/ repl", dict(sampling=2, flags=capi.FLAG_X_TILE_REPLACE | capi.FLAG_X_STEP_RANDOM)),
                ("tile2 warp segments w/ repl", dict(sampling=2, flags=capi.FLAG_X_SEGMENT_RANDOM)),
                ("tile2 scrambled + exch", dict(sampling=2, flags=M | capi.FLAG_EXCH_WRITE))]
with odgi_b200.Engine(g) as e:
    for name, kw in VARIANTS:
        far, loc, rate = [], [], []
        for seed in seeds:
            cd = capi.layout_defaults(g, seed=seed, **kw)
            e.set_coords_2d(X0, Y0)
            st = e.run_2d(cd)
            far.append(e.path_stress(2, 4_000_000, 12345))
            loc.append(e.local_stress(2, 4_000_000, 12345))
            rate.append(st["term_updates"] / st["seconds_iterations"] / 1e9)
        print(f"{name:18s} far mean {np.mean(far):.6g} sd {np.std(far, ddof=1):.3g}  [{' '.join(f'{v:.3g}' for v in far)}]   local mean {np.mean(loc):.4g} sd {np.std(loc, ddof=1):.3g}"
              f"   {np.mean(rate):.1f} G/s", flush=True)
