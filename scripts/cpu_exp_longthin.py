"""CPU experiment (oracle models): which ingredient of the device schedule moves the far-pair stress of a LONG THIN graph
(few haplotypes, path length >> node count scale)?  Scaled-down longthin (same generator, 6 haplotypes): one Hogwild with the
reference's sampler (baseline) vs the sequential tile-ORDER model vs the in-flight staleness model at K = N/32 .. N/2."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odgi_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

n_sites = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
g = synth.generate(n_sites, 6, seed=42)
go = orc.Graph(g.node_len, g.path_first_step, g.step_node, g.step_rev)
X0, Y0 = orc.layout_init(go, 42)
print(f"N={go.N} S={go.S}", flush=True)


def report(tag, xy, t0):
    X, Y = orc.xy_to_XY(xy)
    print(f"{tag:34s} far {orc.path_stress_2d(go, X, Y, 1_000_000, 12345):.6g}  local {orc.local_stress_2d(go, X, Y, 1_000_000, 12345):.5g}   ({time.time() - t0:.0f} s)", flush=True)


for seed in (9399220, 11):
    cfg = orc.default_layout_config(go)
    cfg.seed = seed
    t0 = time.time(); _, xy = orc.layout_2d_f32(go, cfg, orc.XY_to_xy(X0, Y0), n_streams=64); report(f"seed {seed}: one Hogwild, 64 streams", xy, t0)
    t0 = time.time(); xy = orc.XY_to_xy(X0, Y0); orc.run_tile_order(go, cfg, 2048, 2, xy=xy); report(f"seed {seed}: tile-order model", xy, t0)
    for frac in (32, 8, 2):
        K = max(64, go.N // frac)
        t0 = time.time(); xy = orc.XY_to_xy(X0, Y0); orc.run_inflight(go, cfg, K, 2, False, xy=xy); report(f"seed {seed}: in flight N/{frac} (red.add)", xy, t0)
