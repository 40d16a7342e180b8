"""CPU experiment (oracle emulation, round 2): is the replicated + mean-all-reduce multi-GPU schedule quality-safe on graphs
with c4's path depth?  Round 1 measured +50..75 % stress at 8 ranks on DRB1-3123 / chr6.C4 (tens of updates per node and
iteration) but an unchanged stress on c4 itself (760 updates per node and iteration, 90 paths).  Here: the same generator at
30 000 sites with 90 / 16 / 4 paths (same depth as c4 / small / shallow), 8 ranks, far-pair and local stress, against the
single-Hogwild baseline.  Decides the rule PGSGD_MULTI_AUTO uses."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odgi_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

ranks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
PATHS = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [90, 16, 4]
for n_paths in PATHS:
    g = synth.generate(30_000, n_paths, seed=42)
    go = orc.Graph(g.node_len, g.path_first_step, g.step_node, g.step_rev)
    cfg = orc.default_layout_config(go)
    X0, Y0 = orc.layout_init(go, 42)
    upn = cfg.min_term_updates / go.N
    print(f"{n_paths} paths: N={go.N} S={go.S} updates/node/iteration={upn:.0f}  per replica at {ranks} ranks: {upn / ranks:.0f}", flush=True)
    t0 = time.time()
    xy = orc.XY_to_xy(X0, Y0)
    _, xy1 = orc.layout_2d_f32(go, cfg, xy, n_streams=64)
    X, Y = orc.xy_to_XY(xy1)
    base = (orc.path_stress_2d(go, X, Y, 1_000_000, 12345), orc.local_stress_2d(go, X, Y, 1_000_000, 12345))
    print(f"  one Hogwild (64 streams)        far {base[0]:.5g}  local {base[1]:.5g}   ({time.time() - t0:.0f} s)", flush=True)
    for syncs in (1, 4):
        t0 = time.time()
        xym = orc.emulate_multirank_2d_f32(go, cfg, orc.XY_to_xy(X0, Y0), ranks, 8, syncs_per_iter=syncs)
        X, Y = orc.xy_to_XY(xym)
        s = (orc.path_stress_2d(go, X, Y, 1_000_000, 12345), orc.local_stress_2d(go, X, Y, 1_000_000, 12345))
        print(f"  {ranks} replicas, mean x{syncs}/iteration   far {s[0]:.5g} ({s[0] / base[0] - 1:+.1%})  local {s[1]:.5g} ({s[1] / base[1] - 1:+.1%})   ({time.time() - t0:.0f} s)", flush=True)
