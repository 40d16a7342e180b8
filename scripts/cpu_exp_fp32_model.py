"""CPU experiment (oracle): is the GPU runs' offset in final far stress on `mid` 2D the fp32 coordinate format?
The oracle's fp32 model of the device arithmetic (orc_layout_2d_f32: fp32 coordinates and update arithmetic, the reference's
sampler, 6 interleaved streams, sequential) under the same seeds as the fp64 oracle runs of scripts/cpu_exp_seed_spread.py.

  python scripts/cpu_exp_fp32_model.py <graph> <seed> [n_streams=6]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odgi_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

name, seed = sys.argv[1], int(sys.argv[2])
n_streams = int(sys.argv[3]) if len(sys.argv) > 3 else 6
g = synth.preset(name) if name in synth.PRESETS else synth.generate(3_000_000, 6, seed=42)
go = orc.Graph(g.node_len, g.path_first_step, g.step_node, g.step_rev)
cfg = orc.default_layout_config(go)
cfg.seed = seed
X, Y = orc.layout_init(go, 42)
xy = np.empty(4 * go.N, dtype=np.float32)
xy[0::4], xy[1::4], xy[2::4], xy[3::4] = X[0::2], Y[0::2], X[1::2], Y[1::2]
t0 = time.time()
n, xy = orc.layout_2d_f32(go, cfg, xy, n_streams=n_streams)
Xo, Yo = np.empty(2 * go.N), np.empty(2 * go.N)
Xo[0::2], Yo[0::2], Xo[1::2], Yo[1::2] = xy[0::4], xy[1::4], xy[2::4], xy[3::4]
far, loc = orc.path_stress_2d(go, Xo, Yo, 4_000_000, 12345), orc.local_stress_2d(go, Xo, Yo, 4_000_000, 12345)
print(f"{name} fp32 model seed={seed} n_streams={n_streams}: far {far:.6g} local {loc:.5g} updates {n} ({time.time() - t0:.0f} s)", flush=True)
with open(os.path.join(ROOT, ".scratch", "scale_golden", "oracle_fp32_runs.jsonl"), "a") as f:
    f.write(json.dumps({"graph": name, "kind": "layout2d", "model": "fp32", "seed": seed, "n_streams": n_streams, "far": far, "local": loc}) + "\n")
