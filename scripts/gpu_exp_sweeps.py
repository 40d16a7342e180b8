"""GPU experiment (round 2): T terms per staged step and visit (PGSGD_TILE_SWEEPS): a tile is fetched q/T times per iteration
instead of q times — fewer DRAM sectors per update if the sector-rate wall counts the sequential tile stream too.
Throughput early / cooling and the final far / local stress (c4 is reproducible to < 1 % between seeds)."""
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
    for T in (1, 2, 5, 10, 1):
        os.environ["PGSGD_TILE_SWEEPS"] = str(T)
        for seed in ((9399220, 7) if wl != "c4" else (9399220,)):
            e.set_coords_2d(X0, Y0)
            cd = capi.layout_defaults(g, sampling=2, seed=seed)
            e.run_range(cd, 2, 0, 1)
            st = e.run_range(cd, 2, 1, 4)
            e.run_range(cd, 2, 4, 20)
            st2 = e.run_range(cd, 2, 20, 23)
            e.run_range(cd, 2, 23, 30)
            print(f"T={T:2d} seed {seed}: early {st['term_updates'] / st['seconds_iterations'] / 1e9:6.2f} G/s   cooling {st2['term_updates'] / st2['seconds_iterations'] / 1e9:6.2f} G/s"
                  f"   far {e.path_stress(2, 4_000_000, 12345):.6g}  local {e.local_stress(2, 4_000_000, 12345):.5g}  (counted {st['term_updates']})", flush=True)
