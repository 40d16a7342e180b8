"""GPU experiment (round 2): tile visiting order and L2 residency of the pipelined tile kernel.

The r02 ncu captures show the same DRAM sector rate (~69 G 32-byte sectors/s) in the early and in the cooling iterations while
nothing on the SM side is saturated: the wall is the rate of RANDOM DRAM sectors (far partner records + coordinate misses /
write-backs), not bytes.  Variants: random tile bijection (default) vs sweep order (resident CTAs share one window of the step
array -> far partners hit L2), coordinates pinned by an access-policy window.  Throughput early / cooling and the final far
and local stress of a complete default schedule (same initial layout for all)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
g = synth.preset(wl) if wl in synth.PRESETS else synth.generate(3_000_000, 6, seed=42)
X0, Y0 = odgi_b200.layout_init(g, 42)
print(f"workload={wl} N={g.N} S={g.S}", flush=True)
SWEEP, WINDOW, HALF, TMA = capi.FLAG_SWEEP_TILES, capi.FLAG_L2_WINDOW, capi.FLAG_HALF_TILE, capi.FLAG_TMA_STAGING
VARIANTS = [("tile2 random order", 0), ("tile2 sweep", SWEEP), ("tile2 sweep + L2 window", SWEEP | WINDOW), ("tile2 L2 window", WINDOW),
            ("tile2 sweep TMA 2x1024", SWEEP | TMA | HALF)]
with odgi_b200.Engine(g) as e:
    for name, flags in VARIANTS:
        e.set_coords_2d(X0, Y0)
        cd = capi.layout_defaults(g, sampling=2, flags=flags)
        e.run_range(cd, 2, 0, 1)
        st = e.run_range(cd, 2, 1, 4)
        e.run_range(cd, 2, 4, 20)
        st2 = e.run_range(cd, 2, 20, 23)
        e.run_range(cd, 2, 23, 30)
        far, loc = e.path_stress(2, 4_000_000, 12345), e.local_stress(2, 4_000_000, 12345)
        print(f"{name:26s} early {st['term_updates'] / st['seconds_iterations'] / 1e9:6.2f} G/s   cooling "
              f"{st2['term_updates'] / st2['seconds_iterations'] / 1e9:6.2f} G/s   final stress far {far:.6g} local {loc:.5g}", flush=True)
