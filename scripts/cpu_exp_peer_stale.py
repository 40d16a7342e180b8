"""CPU experiment (oracle emulation): multi-GPU peer phase with STALE REMOTE READS — a rank reads nodes it does not own from a
snapshot refreshed R times per iteration, every update still lands on the one true array (orc_peer_stale_2d_f32).  How many
refreshes per iteration keep the final stress?  (Round-2 design input: reading remote partners from a local replica would
halve the fine-grained NVLink operations of the peer phase)  2D, hybrid schedule: iterations < iter_max/3 as
one Hogwild (stand-in for the all-reduce phase), then the stale-read peer phase; several seeds."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odgi_b200 import synth  # noqa: E402
from odgi_b200.arrays import read_arrays  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def graph(name):
    if name in synth.PRESETS:
        g = synth.preset(name)
        return orc.Graph(g.node_len, g.path_first_step, g.step_node, g.step_rev)
    return orc.Graph.from_arrays(read_arrays(os.path.join(ROOT, "tests", "golden", f"{name}.graph.arr.gz")))


ranks = 8
for name in sys.argv[1:] or ["chr6.C4", "small"]:
    go = graph(name)
    print(f"{name}: N={go.N} S={go.S}, {ranks} ranks", flush=True)
    for refreshes in (0, 1, 4, 16, 64):
        vals = []
        for seed in (9399220, 11, 12):
            cfg = orc.default_layout_config(go)
            cfg.seed = seed
            sw = cfg.iter_max // 3
            X0, Y0 = orc.layout_init(go, 42)
            xy = orc.XY_to_xy(X0, Y0)
            st = np.zeros(4 * 64, dtype=np.uint64)
            orc.run_range(go, cfg, 64, cfg.seed, cfg.min_term_updates, 0, sw, 1, xy=xy, rng_state=st)
            if refreshes == 0:      # baseline: fresh reads everywhere
                st = np.zeros(4 * 64, dtype=np.uint64)
                orc.run_range(go, cfg, 64, cfg.seed + 1000, cfg.min_term_updates, sw, cfg.iter_max, 1, xy=xy, rng_state=st)
            else:
                orc.peer_stale_2d_f32(go, cfg, xy, ranks, 8, refreshes, sw, cfg.iter_max)
            vals.append(orc.path_stress_2d(go, *orc.xy_to_XY(xy), 1000000, 12345))
        tag = "fresh reads (baseline)" if refreshes == 0 else f"{refreshes:3d} refreshes/iteration"
        print(f"  {tag:24s} stress {np.mean(vals):.5g}  (runs: {' '.join(f'{v:.5g}' for v in vals)})", flush=True)
