"""GPU diagnostic (round 2): tile vs stream sampling as a function of tile size and of haplotype depth (steps per node).
longthin (6 haplotypes) ends above the reference band with tile sampling and inside it with stream sampling; mid (90) shows no
difference.  Far / local stress over several seeds: tile 1024 / 2048 / 4096 on longthin, and tile vs stream on graphs of the
same generator with 12, 24 and 48 haplotypes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

seeds = [9399220, 1234567, 42, 7, 8, 9, 10, 11]


def run(g, variants):
    X0, Y0 = odgi_b200.layout_init(g, 42)
    with odgi_b200.Engine(g) as e:
        for name, kw in variants:
            far, loc = [], []
            for seed in seeds:
                cd = capi.layout_defaults(g, seed=seed, **kw)
                e.set_coords_2d(X0, Y0)
                e.run_2d(cd)
                far.append(e.path_stress(2, 4_000_000, 12345))
                loc.append(e.local_stress(2, 4_000_000, 12345))
            print(f"  {name:16s} far mean {np.mean(far):.6g} sd {np.std(far, ddof=1):.3g} max {np.max(far):.3g}   local mean {np.mean(loc):.4g} sd {np.std(loc, ddof=1):.3g}", flush=True)


g = synth.generate(3_000_000, 6, seed=42)
print(f"longthin: N={g.N} S={g.S} depth {g.S / g.N:.1f}", flush=True)
run(g, [("tile 1024", dict(sampling=2, flags=capi.FLAG_HALF_TILE)), ("tile 2048", dict(sampling=2)), ("tile 4096", dict(sampling=2, flags=capi.FLAG_BIG_TILE)),
        ("stream", dict(sampling=1))])
for n_sites, n_paths in ((1_500_000, 12), (1_000_000, 24), (500_000, 48)):
    g = synth.generate(n_sites, n_paths, seed=42)
    print(f"{n_sites} sites x {n_paths} haplotypes: N={g.N} S={g.S} depth {g.S / g.N:.1f}", flush=True)
    run(g, [("tile 2048", dict(sampling=2)), ("stream", dict(sampling=1))])
