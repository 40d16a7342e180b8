"""Multi-GPU measurement suite (round 2), ONE torchrun launch for several measurements (8-GPU box time is charged 8x):

  torchrun --nproc-per-node N scripts/gpu_multi_suite.py [c4] [mid] [longthin]

For every workload and every mode (allreduce, hybrid; auto's choice is printed): K = 12 timed iterations [3, 15) by host wall
clock between barriers (max over ranks) and by CUDA events, rank 0's kernel | collective split, then the rest of the schedule
and the final far / local stress (collective readout).  One JSON line per (workload, mode) on rank 0."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def bcast_id():
    obj = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    return obj[0]


def maxr(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sumr(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c4"]
ALL = {"allreduce": capi.MULTI_ALLREDUCE, "hybrid": capi.MULTI_HYBRID, "peer": capi.MULTI_PEER, "auto": capi.MULTI_AUTO, "single": capi.MULTI_SINGLE}
modes = [("allreduce", capi.MULTI_ALLREDUCE), ("hybrid", capi.MULTI_HYBRID)]
if "--peer" in sys.argv:
    modes.append(("peer", capi.MULTI_PEER))
for a in sys.argv[1:]:
    if a.startswith("--modes="):
        modes = [(m, ALL[m]) for m in a[8:].split(",")]
W, K = 3, 12
for name in names:
    g = synth.preset(name) if name in synth.PRESETS else synth.generate(3_000_000, 6, seed=42)
    X0, Y0 = odgi_b200.layout_init(g, 42)
    cfg = capi.layout_defaults(g)
    e = odgi_b200.Engine(g, device=local)
    e.attach_comm(bcast_id(), world, rank)
    e.set_multi_mode(capi.MULTI_AUTO)
    e.set_coords_2d(X0, Y0)
    auto = {capi.MULTI_ALLREDUCE: "allreduce", capi.MULTI_HYBRID: "hybrid", capi.MULTI_PEER: "peer", capi.MULTI_SINGLE: "single"}[e.resolved_multi_mode()]
    e.close()
    for mname, mode in modes:
        e = odgi_b200.Engine(g, device=local)
        e.attach_comm(bcast_id(), world, rank)
        e.set_multi_mode(mode)
        e.set_coords_2d(X0, Y0)
        dist.barrier(); torch.cuda.synchronize()
        e.run_range(cfg, 2, 0, W)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.time()
        st = e.run_range(cfg, 2, W, W + K)
        torch.cuda.synchronize()
        wall = maxr(time.time() - t0)
        dist.barrier()
        dev = maxr(st["seconds_iterations"])
        upd = sumr(st["term_updates"])
        e.run_range(cfg, 2, W + K, 30)
        far, loc = e.path_stress(2, 4_000_000, 12345), e.local_stress(2, 4_000_000, 12345)
        if rank == 0:
            print(json.dumps({"workload": name, "nodes": g.N, "steps": g.S, "n_gpus": world, "mode": mname, "auto_would_pick": auto,
                              "sampling": {1: "stream", 2: "tile"}.get(st["sampling_used"]),
                              "G_updates_per_s_wall": upd / wall / 1e9, "G_updates_per_s_events": upd / dev / 1e9, "wall_s": wall, "events_s": dev,
                              "rank0_kernel_s": st["seconds_kernels"], "rank0_collective_s": st["seconds_collectives"],
                              "stress_far": far, "stress_local": loc}), flush=True)
        e.close()
dist.destroy_process_group()
