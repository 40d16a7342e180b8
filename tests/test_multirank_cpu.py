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
# path-sharded step records (pgsgd_engine_set_shard): every rank holds only ITS paths, dealt out by the product's host helper,
# and does U * S_rank / S updates per iteration on them; replicas are averaged as above
import odgi_b200
arrs = read_arrays(os.path.join(os.environ["PGSGD_ROOT"], "tests/golden/DRB1-3123.graph.arr.gz"))
gd = odgi_b200.graph_from_arrays(arrs)
mine = odgi_b200.shard_paths(gd, world, rank)
g_mine = orc.Graph(mine.node_len, mine.path_first_step, mine.step_node, mine.step_rev)
xy = orc.XY_to_xy(X0, Y0)
state = np.zeros(4 * n_streams, dtype=np.uint64)
share = cfg.min_term_updates * mine.S // gd.S                  # pgsgd_capi.cu: U_rank of a sharded engine
total = 0
for it in range(cfg.iter_max):
    total += orc.run_range(g_mine, cfg, n_streams, cfg.seed + rank * n_streams, share, it, it + 1, 1, xy=xy, rng_state=state)
    t = torch.from_numpy(xy.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    xy = (t.numpy() / np.float32(world)).astype(np.float32)
shards = []
for r in range(world):
    s = odgi_b200.shard_paths(gd, world, r)
    shards.append(orc.Graph(s.node_len, s.path_first_step, s.step_node, s.step_rev))
ref = orc.emulate_sharded_2d_f32(shards, gd.S, cfg, orc.XY_to_xy(X0, Y0), n_streams)
cnt = torch.tensor([total]); dist.all_reduce(cnt)
steps = torch.tensor([mine.S]); dist.all_reduce(steps)
res["sharded"] = {"equal": bool(np.array_equal(xy, ref)), "updates": int(cnt.item()), "steps": int(steps.item()),
                  "expected_updates": int(sum(cfg.min_term_updates * s.S // gd.S for s in shards) * cfg.iter_max)}
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
    sh = res["sharded"]
    assert sh["equal"] and sh["steps"] == 35059 and sh["updates"] == sh["expected_updates"] and 5 * 20001 - 5 * 2 < sh["updates"] <= 5 * 20001, res


def test_path_assignment_is_balanced_and_complete():
    import numpy as np
    import odgi_b200
    from odgi_b200.arrays import read_arrays
    g = odgi_b200.graph_from_arrays(read_arrays(os.path.join(ROOT, "tests", "golden", "chr6.C4.graph.arr.gz")))
    counts = np.diff(g.path_first_step)
    for n in (1, 2, 3, 8):
        owner = odgi_b200.assign_paths(counts, n)
        loads = np.array([counts[owner == r].sum() for r in range(n)])
        assert loads.sum() == g.S and loads.max() - loads.min() <= counts.max()      # greedy LPT bound
        seen = 0
        for r in range(n):
            s = odgi_b200.shard_paths(g, n, r)
            assert s.S == loads[r] and s.P == int((owner == r).sum()) and s.N == g.N
            # a shard's paths keep their own steps, orientation and positions
            p0 = int(np.nonzero(owner == r)[0][0])
            lo, hi = int(g.path_first_step[p0]), int(g.path_first_step[p0 + 1])
            assert np.array_equal(s.step_node[: hi - lo], g.step_node[lo:hi]) and np.array_equal(s.step_pos[: hi - lo], g.step_pos[lo:hi])
            seen += s.S
        assert seen == g.S


def test_rank_local_generator_shards_one_well_defined_graph():
    """synth.generate_sharded (BASELINE config 5's generator): the union of the ranks' shards is the same graph whatever the
    number of ranks; every rank carries the whole node table; paths are dealt out p = rank (mod n_ranks)."""
    import numpy as np
    from odgi_b200 import synth
    kw = dict(seed=5, inv_per_mbp=20.0, dup_per_mbp=10.0)   # enough structural events at this size to exercise them
    whole, ids = synth.generate_sharded(30_000, 7, 0, 1, **kw)
    assert ids == list(range(7)) and whole.P == 7
    wf = whole.path_first_step.astype(np.int64)
    for n_ranks in (2, 3, 8):
        seen, steps = [], 0
        for r in range(n_ranks):
            g, mine = synth.generate_sharded(30_000, 7, r, n_ranks, **kw)
            assert mine == list(range(r, 7, n_ranks)) and np.array_equal(g.node_len, whole.node_len)
            f = g.path_first_step.astype(np.int64)
            for j, p in enumerate(mine):
                assert np.array_equal(g.step_node[f[j]:f[j + 1]], whole.step_node[wf[p]:wf[p + 1]])
                assert np.array_equal(g.step_rev[f[j]:f[j + 1]], whole.step_rev[wf[p]:wf[p + 1]])
            seen += mine
            steps += g.S
        assert sorted(seen) == list(range(7)) and steps == whole.S
    assert whole.step_node.max() < whole.N and whole.step_rev.sum() > 0 and whole.S > 7 * 30_000   # inversions and duplications present
