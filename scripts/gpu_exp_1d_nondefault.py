"""1D with a non-default schedule (30 iterations of 10*S updates) on small graphs: stream vs tile vs oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import odgi_b200
from odgi_b200 import capi
from odgi_b200.arrays import read_arrays
from oracle import oracle as orc
for name in ("LPA", "DRB1-3123"):
    a = read_arrays(os.path.join(ROOT, f"tests/golden/{name}.graph.arr.gz"))
    gd, go = odgi_b200.graph_from_arrays(a), orc.Graph.from_arrays(a)
    kw = dict(iter_max=30, min_term_updates=10 * gd.S)
    co = orc.default_sort_config(go, **kw)
    n, x = orc.sort_1d(go, co, orc.sort_init(go), n_streams=64)
    print(name, "oracle 64 streams", orc.path_stress_1d(go, x, 1000000, 12345), flush=True)
    for sampling, nm in ((1, "stream"), (2, "tile")):
        for seed in (1, 2):
            x, st = odgi_b200.sort_1d(gd, capi.sort_defaults(gd, sampling=sampling, seed=seed, **kw))
            print(name, nm, seed, orc.path_stress_1d(go, x, 1000000, 12345), st["term_updates"], flush=True)
