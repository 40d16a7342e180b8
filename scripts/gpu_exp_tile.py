"""GPU experiment: stream vs tile sampling throughput on a synthetic graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200
from odgi_b200 import capi, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "mid"
g = synth.preset(wl)
X0, Y0 = odgi_b200.layout_init(g, 42)
print(f"workload={wl} N={g.N} S={g.S}", flush=True)
with odgi_b200.Engine(g) as e:
    e.set_coords_2d(X0, Y0)
    for sampling, name, flags in ((1, "stream", 0), (2, "tile", 0), (2, "tileTMA", 8)):
        for batch in ((1,) if sampling == 1 else ((1, 2, 4) if flags == 0 else (4,))):
            cd = capi.layout_defaults(g, batch=batch, sampling=sampling, flags=flags)
            e.run_range(cd, 2, 0, 1)
            st = e.run_range(cd, 2, 1, 4)
            st2 = e.run_range(cd, 2, 20, 23)
            print(f"{name:7s} batch={batch}  early {st['term_updates']/st['seconds_iterations']/1e9:6.2f} G/s   cooling {st2['term_updates']/st2['seconds_iterations']/1e9:6.2f} G/s  (counted {st['term_updates']})", flush=True)
