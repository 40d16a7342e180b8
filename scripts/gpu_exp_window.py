"""GPU experiment (round 2): window order of the tile visits (PGSGD_FLAG_WINDOW_TILES).

DESIGN.md 3.6: beyond the partner-record gather the cost of a term is coordinate sectors that fell out of L2 (c4: 0.3-0.4 of the
1.64 DRAM sectors per update; c4x and larger: most of them).  In window order the tiles of all paths over one stretch of the
node order are visited together (C of them at a time), so that stretch's coordinates are fetched once and reused ~depth times.
Throughput early / cooling and the final far / local stress of complete default schedules, against the random tile bijection."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
seeds = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [9399220]
g = synth.preset(wl) if wl in synth.PRESETS else synth.generate(3_000_000, 6, seed=42)
X0, Y0 = odgi_b200.layout_init(g, 42)
print(f"workload={wl} N={g.N} S={g.S} seeds={seeds}", flush=True)
VARIANTS = [("random bijection (default)", 0, None), ("window order C=2", capi.FLAG_WINDOW_TILES, "2"), ("window order C=3", capi.FLAG_WINDOW_TILES, "3"),
            ("window order C=6", capi.FLAG_WINDOW_TILES, "6"), ("window order C=12", capi.FLAG_WINDOW_TILES, "12")]
with odgi_b200.Engine(g) as e:
    for name, flags, c in VARIANTS:
        if c is not None:
            os.environ["PGSGD_WINDOW_C"] = c
        fars, locs, early, cool = [], [], [], []
        for seed in seeds:
            e.set_coords_2d(X0, Y0)
            cd = capi.layout_defaults(g, sampling=2, flags=flags, seed=seed)
            e.run_range(cd, 2, 0, 1)
            st = e.run_range(cd, 2, 1, 4)
            e.run_range(cd, 2, 4, 20)
            st2 = e.run_range(cd, 2, 20, 23)
            e.run_range(cd, 2, 23, 30)
            fars.append(e.path_stress(2, 4_000_000, 12345)); locs.append(e.local_stress(2, 4_000_000, 12345))
            early.append(st["term_updates"] / st["seconds_iterations"] / 1e9); cool.append(st2["term_updates"] / st2["seconds_iterations"] / 1e9)
        print(f"{name:28s} early {sum(early) / len(early):6.2f} G/s   cooling {sum(cool) / len(cool):6.2f} G/s   final stress far "
              f"[{' '.join(f'{x:.6g}' for x in fars)}]  local [{' '.join(f'{x:.5g}' for x in locs)}]", flush=True)
