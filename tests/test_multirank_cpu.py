"""The N > 1 schedule on CPU, world_size 2 over gloo: every rank runs its share of each iteration's updates on its own
replica (the oracle stands in for the kernels), replicas are combined with an all-reduce — the same split, seeds and
combine the NCCL path uses (pgsgd_capi.cu run_engine).  Checked against a single-process emulation bit for bit."""
import os
import subprocess
import sys

import json
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, os.environ["PGSGD_ROOT"])
import torch, torch.distributed as dist
from odgi_b200.arrays import read_arrays
from oracle import oracle as orc
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
g = orc.Graph.from_arrays(read_arrays(os.path.join(os.environ["PGSGD_ROOT"], "tests/golden/DRB1-3123.graph.arr.gz")))
cfg = orc.default_layout_config(g, iter_max=5, min_term_updates=20001)
X0, Y0 = orc.layout_init(g, 42)
n_streams = 4
res = {}
for sum_deltas in (False, True):
    xy = orc.XY_to_xy(X0, Y0)
    state = np.zeros(4 * n_streams, dtype=np.uint64)
    U = cfg.min_term_updates
    share = U // world + (1 if rank < U % world else 0)     # pgsgd_capi.cu: U_rank
    total = 0
    for it in range(cfg.iter_max):
        prev = xy.copy()
        total += orc.run_range(g, cfg, n_streams, cfg.seed + rank * n_streams, share, it, it + 1, 1, xy=xy, rng_state=state)
        t = torch.from_numpy(xy - prev if sum_deltas else xy.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        xy = (prev + t.numpy()) if sum_deltas else (t.numpy() / np.float32(world)).astype(np.float32)
    ref = orc.emulate_multirank_2d_f32(g, cfg, orc.XY_to_xy(X0, Y0), world, n_streams, sum_deltas=sum_deltas)
    cnt = torch.tensor([total]); dist.all_reduce(cnt)
    res["sum" if sum_deltas else "avg"] = {"equal": bool(np.array_equal(xy, ref)), "updates": int(cnt.item())}
if rank == 0:
    print("RESULT " + json.dumps(res))
dist.destroy_process_group()
'''


def test_two_rank_schedule_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, PGSGD_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29513", str(script)], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1][7:])
    for tag in ("avg", "sum"):
        assert res[tag]["updates"] == 5 * 20001, res
        assert res[tag]["equal"], res
