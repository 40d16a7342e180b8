"""forced tile sampling on the small fixture graphs: stress vs reference band for several launch shapes"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import odgi_b200
from odgi_b200 import capi
from odgi_b200.arrays import read_arrays
from oracle import oracle as orc
band = json.load(open(os.path.join(ROOT, "tests/golden/stress_reference.json")))
for name, dims in (("LPA", 1), ("DRB1-3123", 1), ("DRB1-3123", 2), ("chr6.C4", 2)):
    a = read_arrays(os.path.join(ROOT, f"tests/golden/{name}.graph.arr.gz"))
    gd, go = odgi_b200.graph_from_arrays(a), orc.Graph.from_arrays(a)
    ref = band[f"{name}.{'layout2d' if dims == 2 else 'sort1d'}"]
    X0, Y0 = orc.layout_init(go, 42)
    for batch in (1, 2):
        for ctas in (0, 1, 2, 4, 8, 16):
            vals = []
            for seed in (1, 2):
                mk = capi.layout_defaults if dims == 2 else capi.sort_defaults
                cd = mk(gd, sampling=2, batch=batch, n_streams=ctas * 256, seed=seed)
                if dims == 2:
                    X, Y, st = odgi_b200.layout_2d(gd, cd, X0, Y0)
                    vals.append(orc.path_stress_2d(go, X, Y, ref["n_pairs"], ref["seed"]))
                else:
                    x, st = odgi_b200.sort_1d(gd, cd)
                    vals.append(orc.path_stress_1d(go, x, ref["n_pairs"], ref["seed"]))
            print(f"{name} {dims}D N={gd.N} batch={batch} CTAs={'auto' if ctas == 0 else ctas}: " + " ".join(f"{v:.5g} ({(v-ref['mean'])/ref['mean']*100:+.1f}%)" for v in vals), flush=True)
