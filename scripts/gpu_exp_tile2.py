"""GPU experiment (round 2): legacy tile kernel vs the pipelined tile kernel and its staging / tile-size variants.

  python scripts/gpu_exp_tile2.py [workload] [--quality]

Throughput of iterations 1-3 (half of the partners uniform over the path) and 20-22 (cooling: all Zipf) per variant on one
resident engine; with --quality also the final sampled path stress of a complete default schedule for the legacy and the
default pipelined kernel (same initial layout)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
wl = args[0] if args else "c4"
g = synth.preset(wl)
X0, Y0 = odgi_b200.layout_init(g, 42)
print(f"workload={wl} N={g.N} S={g.S}", flush=True)
LEGACY, TMA, HALF, BIG = 32, 8, 64, 128
VARIANTS = [("legacy tile (r01 default)", LEGACY), ("tile2 LDG+STS 2048", 0), ("tile2 LDG+STS 1024", HALF), ("tile2 LDG+STS 4096", BIG),
            ("tile2 TMA 2x1024", TMA | HALF), ("tile2 TMA 2x2048", TMA), ("legacy TMA 2x2048", LEGACY | TMA)]
with odgi_b200.Engine(g) as e:
    for name, flags in VARIANTS:
        e.set_coords_2d(X0, Y0)
        cd = capi.layout_defaults(g, sampling=2, flags=flags)
        try:
            e.run_range(cd, 2, 0, 1)
            st = e.run_range(cd, 2, 1, 4)
            st2 = e.run_range(cd, 2, 20, 23)
        except Exception as ex:  # a variant that does not fit must not lose the others
            print(f"{name:28s} FAILED: {ex}", flush=True)
            continue
        print(f"{name:28s} early {st['term_updates'] / st['seconds_iterations'] / 1e9:6.2f} G/s   cooling "
              f"{st2['term_updates'] / st2['seconds_iterations'] / 1e9:6.2f} G/s  (counted {st['term_updates']}, flags_used {st['flags_used']})", flush=True)
    if "--quality" in sys.argv:
        for name, flags in (("legacy tile", LEGACY), ("tile2 default", 0), ("tile2 TMA 2x1024", TMA | HALF)):
            e.set_coords_2d(X0, Y0)
            cd = capi.layout_defaults(g, sampling=2, flags=flags)
            s0 = e.path_stress(2, 4_000_000, 12345)
            st = e.run_range(cd, 2, 0, 30)
            s1 = e.path_stress(2, 4_000_000, 12345)
            print(f"{name:20s} full schedule: {st['term_updates'] / st['seconds_iterations'] / 1e9:6.2f} G/s  stress {s0:.4f} -> {s1:.6f}", flush=True)
