"""GPU experiment (round 2): what the with-replacement variants of the tile sampling cost on c4 (early / cooling G updates/s)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
g = synth.preset(wl)
X0, Y0 = odgi_b200.layout_init(g, 42)
T, S, G, M = capi.FLAG_X_TILE_REPLACE, capi.FLAG_X_STEP_RANDOM, capi.FLAG_X_SEGMENT_RANDOM, capi.FLAG_X_STEP_SCRAMBLE
print(f"workload={wl} N={g.N} S={g.S}", flush=True)
with odgi_b200.Engine(g) as e:
    for name, flags in (("default (bijections)", 0), ("tiles w/ repl", T), ("tiles + steps w/ repl", T | S), ("tiles + warp segments w/ repl", T | G),
                        ("warp segments w/ repl", G), ("scrambled lanes", M), ("scrambled pairs", M | capi.FLAG_X_SCRAMBLE_PAIRS), ("scrambled quads", M | capi.FLAG_X_SCRAMBLE_QUADS), ("steps w/ repl", S)):
        e.set_coords_2d(X0, Y0)
        cd = capi.layout_defaults(g, sampling=capi.SAMPLING_TILE, flags=flags)
        e.run_range(cd, 2, 0, 1)
        st = e.run_range(cd, 2, 1, 4)
        e.run_range(cd, 2, 4, 20)
        st2 = e.run_range(cd, 2, 20, 23)
        e.run_range(cd, 2, 23, 30)
        print(f"{name:32s} early {st['term_updates'] / st['seconds_iterations'] / 1e9:6.2f} G/s   cooling {st2['term_updates'] / st2['seconds_iterations'] / 1e9:6.2f} G/s   "
              f"final stress far {e.path_stress(2, 4_000_000, 12345):.6g} local {e.local_stress(2, 4_000_000, 12345):.5g}", flush=True)
