"""GPU experiment: stress of default runs vs number of concurrent worker streams / batch / update mode."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200
from odgi_b200 import capi
from odgi_b200.arrays import read_arrays
from oracle import oracle as orc

band = json.load(open(os.path.join(ROOT, "tests/golden/stress_reference.json")))
for name in ("DRB1-3123", "chr6.C4"):
    a = read_arrays(os.path.join(ROOT, f"tests/golden/{name}.graph.arr.gz"))
    gd, go = odgi_b200.graph_from_arrays(a), orc.Graph.from_arrays(a)
    X0, Y0 = orc.layout_init(go, 42)
    ref = band[f"{name}.layout2d"]
    print(f"== {name} 2D  N={gd.N} S={gd.S} reference stress {ref['mean']:.5f} +- {ref['sd']:.5f}")
    with odgi_b200.Engine(gd) as e:
        for flags in (0, 1):
            for batch in (1, 4):
                for ns in (32, 64, 128, 256, 512, 1024, 4096, 0):
                    cd = capi.layout_defaults(gd, n_streams=ns, batch=batch, flags=flags)
                    e.set_coords_2d(X0, Y0)
                    st = e.run_2d(cd)
                    X, Y = e.get_coords_2d()
                    s = orc.path_stress_2d(go, X, Y, ref["n_pairs"], ref["seed"])
                    print(f"flags={flags} batch={batch} streams={ns:5d}  stress={s:.5f}  rel={(s-ref['mean'])/ref['mean']*100:+.2f}%  {st['term_updates']/st['seconds_iterations']/1e6:8.1f} M/s", flush=True)
for name in ("LPA", "DRB1-3123"):
    a = read_arrays(os.path.join(ROOT, f"tests/golden/{name}.graph.arr.gz"))
    gd, go = odgi_b200.graph_from_arrays(a), orc.Graph.from_arrays(a)
    ref = band[f"{name}.sort1d"]
    print(f"== {name} 1D  N={gd.N} S={gd.S} reference stress {ref['mean']:.5f} +- {ref['sd']:.5f}")
    with odgi_b200.Engine(gd) as e:
        for batch in (1, 4):
            for ns in (32, 128, 512, 0):
                cd = capi.sort_defaults(gd, n_streams=ns, batch=batch)
                e.set_coords_1d(None)
                st = e.run_1d(cd)
                x = e.get_coords_1d()
                s = orc.path_stress_1d(go, x, ref["n_pairs"], ref["seed"])
                print(f"batch={batch} streams={ns:5d}  stress={s:.5f}  rel={(s-ref['mean'])/ref['mean']*100:+.2f}%  {st['term_updates']/st['seconds_iterations']/1e6:8.1f} M/s", flush=True)
