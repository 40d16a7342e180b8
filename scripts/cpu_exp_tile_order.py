"""CPU experiment: what the blocked ORDER of tile sampling does to the result, isolated from concurrency and from the fp32
sampler (orc_run_tile_order: every step is a first step floor(U/S) times per iteration, tiles in a fresh random order per
pass, exact partner law, sequential) against the reference's i.i.d. uniform first pick (single stream), full default schedules.
Final sampled path stress; results in DESIGN.md section 5."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odgi_b200 import synth  # noqa: E402
from odgi_b200.arrays import read_arrays  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def graph(name):
    if name in synth.PRESETS or name.startswith("synth:"):
        g = synth.preset(name) if name in synth.PRESETS else synth.generate(*[int(v) for v in name[6:].split("x")], seed=42)
        return orc.Graph(g.node_len, g.path_first_step, g.step_node, g.step_rev)
    return orc.Graph.from_arrays(read_arrays(os.path.join(ROOT, "tests", "golden", f"{name}.graph.arr.gz")))


for name in sys.argv[1:] or ["DRB1-3123", "chr6.C4", "LPA", "small"]:
    go = graph(name)
    print(f"{name}: N={go.N} S={go.S}", flush=True)
    for dims in (2, 1):
        cfg = orc.default_layout_config(go) if dims == 2 else orc.default_sort_config(go)
        res = {}
        for tag in ("uniform", "tiles"):
            t0 = time.time()
            if dims == 2:
                X0, Y0 = orc.layout_init(go, 42)
                xy = orc.XY_to_xy(X0, Y0)
                if tag == "uniform":
                    orc.run_range(go, cfg, 1, cfg.seed, cfg.min_term_updates, 0, cfg.iter_max, 1, xy=xy)
                else:
                    orc.run_tile_order(go, cfg, 2048, 2, xy=xy)
                X, Y = orc.xy_to_XY(xy)
                res[tag] = orc.path_stress_2d(go, X, Y, 1000000, 12345)
            else:
                x = orc.sort_init(go)
                if tag == "uniform":
                    orc.run_range(go, cfg, 1, cfg.seed, cfg.min_term_updates, 0, cfg.iter_max + 1, 2, X=x)
                else:
                    orc.run_tile_order(go, cfg, 2048, 1, X=x)
                res[tag] = orc.path_stress_1d(go, x, 1000000, 12345)
            res[tag + "_s"] = time.time() - t0
        print(f"  {dims}D  uniform first pick {res['uniform']:.5g}   tile order {res['tiles']:.5g}   ({100 * (res['tiles'] / res['uniform'] - 1):+.2f} %)"
              f"   [{res['uniform_s']:.0f} s / {res['tiles_s']:.0f} s]", flush=True)
