"""CPU experiment (oracle = the pinned restatement of the reference): how much of the final stress of a full default run is the
choice of the worker-stream SEEDS?  The reference hard-codes its seeds (9399220 + thread id, path_sgd.cpp / path_sgd_layout.cpp:168),
so its own run-to-run spread (tests/golden/stress_reference_scale.json) is thread-timing noise around ONE seed set.  The oracle
runs the same schedule with n_streams worker streams interleaved term by term (sequential, deterministic) for any seed.

  python scripts/cpu_exp_seed_spread.py <graph> <dims> <seed> [n_streams=6]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odgi_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

name, dims, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
n_streams = int(sys.argv[4]) if len(sys.argv) > 4 else 6
g = synth.preset(name) if name in synth.PRESETS else synth.generate(3_000_000, 6, seed=42)
go = orc.Graph(g.node_len, g.path_first_step, g.step_node, g.step_rev)
t0 = time.time()
if dims == 1:
    cfg = orc.default_sort_config(go)
    cfg.seed = seed
    n, x = orc.sort_1d(go, cfg, orc.sort_init(go), n_streams=n_streams)
    far, loc = orc.path_stress_1d(go, x, 4_000_000, 12345), orc.local_stress_1d(go, x, 4_000_000, 12345)
else:
    cfg = orc.default_layout_config(go)
    cfg.seed = seed
    X, Y = orc.layout_init(go, 42)
    n, X, Y = orc.layout_2d(go, cfg, X, Y, n_streams=n_streams)
    far, loc = orc.path_stress_2d(go, X, Y, 4_000_000, 12345), orc.local_stress_2d(go, X, Y, 4_000_000, 12345)
print(f"{name} dims={dims} seed={seed} n_streams={n_streams}: far {far:.6g} local {loc:.5g} updates {n} ({time.time() - t0:.0f} s)", flush=True)
import json  # noqa: E402
os.makedirs(os.path.join(ROOT, ".scratch", "scale_golden"), exist_ok=True)
with open(os.path.join(ROOT, ".scratch", "scale_golden", "oracle_runs.jsonl"), "a") as f:   # picked up by make_scale_golden.py bands
    f.write(json.dumps({"graph": name, "kind": "layout2d" if dims == 2 else "sort1d", "seed": seed, "n_streams": n_streams, "far": far, "local": loc}) + "\n")
