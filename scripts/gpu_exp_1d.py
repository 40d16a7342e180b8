"""GPU experiment: 1D PG-SGD (odgi sort -Y) throughput on a synthetic graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200
from odgi_b200 import capi, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "mid"
g = synth.preset(wl)
print(f"workload={wl} N={g.N} S={g.S} max_path_bp={g.max_path_bp}", flush=True)
with odgi_b200.Engine(g) as e:
    for sampling, name in ((1, "stream"), (2, "tile")):
        cd = capi.sort_defaults(g, sampling=sampling, iter_max=30, min_term_updates=10 * g.S)
        e.set_coords_1d(None)
        s0 = e.path_stress(1, 1_000_000, 5)
        e.run_range(cd, 1, 0, 1)
        st = e.run_range(cd, 1, 1, 6)
        st2 = e.run_range(cd, 1, 20, 25)
        e.run_range(cd, 1, 25, 31)
        print(f"1D {name:6s} early {st['term_updates']/st['seconds_iterations']/1e9:6.2f} G/s  cooling {st2['term_updates']/st2['seconds_iterations']/1e9:6.2f} G/s  stress {s0:.4g} -> {e.path_stress(1, 1_000_000, 5):.4g}", flush=True)
