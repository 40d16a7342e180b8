"""GPU experiment (round 2, first call): 1D PG-SGD on a synthetic graph, FULL default schedule, stream vs tile sampling.

profiles/r01_1d_throughput_c4.log compared the two samplings on a truncated schedule (iterations 6..19 skipped to save GPU
time), where the final stress is dominated by the transient; this runs every iteration and reports stress along the way."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "mid"
g = synth.preset(wl)
print(f"workload={wl} N={g.N} S={g.S} max_path_bp={g.max_path_bp}", flush=True)
with odgi_b200.Engine(g) as e:
    # the 1D final stress varies by ~30 % run to run under the reference's own law: several seeds per sampling
    for sampling, name, seed in [(s, n, sd) for sd in (9399220, 1, 2) for s, n in ((capi.SAMPLING_STREAM, "stream"), (capi.SAMPLING_TILE, "tile"))]:
        cd = capi.sort_defaults(g, sampling=sampling)
        cd.seed = seed
        n_iters = cd.iter_max + 1
        e.set_coords_1d(None)
        line = [f"{e.path_stress(1, 1_000_000, 5):.4g}"]
        secs = upd = 0.0
        for lo in range(0, n_iters, 5):
            st = e.run_range(cd, 1, lo, min(lo + 5, n_iters))
            secs += st["seconds_iterations"]
            upd += st["term_updates"]
            line.append(f"{e.path_stress(1, 1_000_000, 5):.4g}")
        print(f"1D {name:6s} seed {seed:8d} {upd / secs / 1e9:6.2f} G updates/s  stress every 5 iterations: {' '.join(line)}", flush=True)
