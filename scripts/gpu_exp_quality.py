"""GPU experiment: final stress of tile vs stream sampling at scale, 2D and 1D (default schedules)."""
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
    for sampling, name in ((1, "stream"), (2, "tile")):
        for seed in (9399220, 77):
            cd = capi.layout_defaults(g, sampling=sampling, seed=seed)
            e.set_coords_2d(X0, Y0)
            s0 = e.path_stress(2, 2_000_000, 5)
            st = e.run_2d(cd)
            print(f"2D {name:6s} seed {seed}: stress {s0:.5g} -> {e.path_stress(2, 2_000_000, 5):.6g}   {st['term_updates']/st['seconds_iterations']/1e9:.1f} G/s", flush=True)
    for sampling, name in ((1, "stream"), (2, "tile")):
        for seed in (9399220, 77):
            cd = capi.sort_defaults(g, sampling=sampling, seed=seed)
            e.set_coords_1d(None)
            s0 = e.path_stress(1, 2_000_000, 5)
            st = e.run_1d(cd)
            print(f"1D {name:6s} seed {seed} (default 101 x 1*S): stress {s0:.5g} -> {e.path_stress(1, 2_000_000, 5):.6g}   {st['term_updates']/st['seconds_iterations']/1e9:.1f} G/s", flush=True)
