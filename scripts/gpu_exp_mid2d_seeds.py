"""GPU experiment (round 2): distribution of the final far / local stress of complete default 2D runs on `mid` over worker-stream
seeds, tile sampling (AUTO's choice, the bench kernel) and stream sampling (reference-exact sampler), same initial layout as the
reference runs of tests/golden/stress_reference_scale.json.  How wide is the seed spread, and is there a bias between the samplers?"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = synth.preset("mid")
X0, Y0 = odgi_b200.layout_init(g, 42)
seeds = [9399220, 1234567, 42] + [1000003 * (k + 1) for k in range(n_seeds - 3)]
X = capi.FLAG_X_TILE_REPLACE, capi.FLAG_X_STEP_RANDOM
VARIANTS = [("tile", capi.SAMPLING_TILE, 0), ("stream", capi.SAMPLING_STREAM, 0),
            ("tile, tiles w/ replacement", capi.SAMPLING_TILE, X[0]), ("tile, steps w/ replacement", capi.SAMPLING_TILE, X[1]),
            ("tile, both w/ replacement", capi.SAMPLING_TILE, X[0] | X[1]),
            ("tile, warp segments w/ repl", capi.SAMPLING_TILE, capi.FLAG_X_SEGMENT_RANDOM),
            ("tile, tiles+segments w/ repl", capi.SAMPLING_TILE, X[0] | capi.FLAG_X_SEGMENT_RANDOM),
            ("tile, scrambled lanes", capi.SAMPLING_TILE, capi.FLAG_X_STEP_SCRAMBLE),
            ("tile, tiles repl + scrambled", capi.SAMPLING_TILE, X[0] | capi.FLAG_X_STEP_SCRAMBLE),
            ("tile, scrambled pairs", capi.SAMPLING_TILE, capi.FLAG_X_STEP_SCRAMBLE | capi.FLAG_X_SCRAMBLE_PAIRS),
            ("tile, scrambled quads", capi.SAMPLING_TILE, capi.FLAG_X_STEP_SCRAMBLE | capi.FLAG_X_SCRAMBLE_QUADS),
            ("tile 1024", capi.SAMPLING_TILE, capi.FLAG_HALF_TILE), ("tile 4096", capi.SAMPLING_TILE, capi.FLAG_BIG_TILE),
            ("tile legacy kernel", capi.SAMPLING_TILE, capi.FLAG_LEGACY_TILE), ("tile exchange write", capi.SAMPLING_TILE, capi.FLAG_EXCH_WRITE),
            ("tile window order", capi.SAMPLING_TILE, capi.FLAG_WINDOW_TILES), ("tile sweep order", capi.SAMPLING_TILE, capi.FLAG_SWEEP_TILES)]
if len(sys.argv) > 2:
    VARIANTS = [v for v in VARIANTS if v[0] in sys.argv[2].split(";")]
if len(sys.argv) > 3:   # "<ctas>": the same variants with that many resident CTAs (terms in flight: is the offset to the CPU runs Hogwild staleness?)
    VARIANTS = [(f"{n}, {sys.argv[3]} CTAs", sm, fl) for n, sm, fl in VARIANTS]
extra = dict(n_streams=int(sys.argv[3]) * 256) if len(sys.argv) > 3 else {}
with odgi_b200.Engine(g) as e:
    for name, sampling, flags in VARIANTS:
        far, loc = [], []
        for seed in seeds:
            e.set_coords_2d(X0, Y0)
            e.run_2d(capi.layout_defaults(g, seed=seed, sampling=sampling, flags=flags, **extra))
            far.append(e.path_stress(2, 4_000_000, 12345)); loc.append(e.local_stress(2, 4_000_000, 12345))
        far, loc = np.array(far), np.array(loc)
        print(f"{name:28s} far mean {far.mean():.6f} sd {far.std(ddof=1):.6f} se {far.std(ddof=1) / np.sqrt(len(far)):.6f} median {np.median(far):.6f}   "
              f"local mean {loc.mean():.1f} sd {loc.std(ddof=1):.1f}", flush=True)
