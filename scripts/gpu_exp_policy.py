"""GPU experiment (round 2): L2 eviction policy of the coordinate LOADS of the pipelined tile kernel (the reds keep evict_last)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
g = synth.preset(wl)
X0, Y0 = odgi_b200.layout_init(g, 42)
print(f"workload={wl} N={g.N} S={g.S}", flush=True)
with odgi_b200.Engine(g) as e:
    for name, flags in (("loads evict_last (default)", 0), ("loads evict_first", 1024), ("loads evict_normal", 2048), ("loads evict_last (again)", 0)):
        e.set_coords_2d(X0, Y0)
        cd = capi.layout_defaults(g, sampling=2, flags=flags)
        e.run_range(cd, 2, 0, 1)
        st = e.run_range(cd, 2, 1, 4)
        e.run_range(cd, 2, 4, 20)
        st2 = e.run_range(cd, 2, 20, 23)
        e.run_range(cd, 2, 23, 30)
        print(f"{name:28s} early {st['term_updates'] / st['seconds_iterations'] / 1e9:6.2f} G/s   cooling {st2['term_updates'] / st2['seconds_iterations'] / 1e9:6.2f} G/s"
              f"   far {e.path_stress(2, 4_000_000, 12345):.6g}", flush=True)
