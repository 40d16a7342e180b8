#!/usr/bin/env python
"""bench.py — M node-pair SGD updates/s of the 2D PG-SGD hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|mid|small|chr6.C4|DRB1-3123] [--impl ours|reference]

A "step" is one cooling-schedule iteration of `odgi layout` = ONE persistent-grid kernel launch that
performs min_term_updates = 10 * S term updates (layout_main.cpp:258).  Default workload "c4": the synthetic
90-haplotype chr6-MHC-scale graph of BASELINE config 4 (~5.5e6 nodes, ~4.2e8 path steps; odgi_b200/synth.py,
seed 42) — the configuration north_star quotes the metric on and the largest single-GPU one; its 6.7 GB of
step records do not fit the 126 MB L2, so every timed step reads its inputs from HBM (no flush needed).

value  : whole-job updates/s with the graph + coordinates resident in HBM; device time from CUDA events on the
         stream the kernels run on (pgsgd_stats.seconds_iterations), max over ranks.
e2e    : the same metric through the reference-facing C-ABI sequence with HOST buffers (what odgi's shim calls):
         flatten-to-device upload of the graph from pinned host memory, coordinate upload, the same K steps,
         coordinate download — host wall clock, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_UPDATE_2D = 80  # algorithmic bytes per 2D term update


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def pinned_like(a: np.ndarray) -> np.ndarray:
    """copy into pinned host memory (torch is only the allocator here)"""
    import torch
    t = torch.empty(a.shape, dtype=getattr(torch, {"uint8": "uint8", "uint32": "int32", "uint64": "int64", "float64": "float64",
                                                  "float32": "float32"}[a.dtype.name]), pin_memory=True)
    v = t.numpy().view(a.dtype)
    v[...] = a
    return v, t  # the caller keeps t alive


def make_workload(name: str):
    import odgi_b200
    from odgi_b200 import synth
    if name in synth.PRESETS:
        g = synth.preset(name, seed=42)
        desc = f"synthetic '{name}' (odgi_b200/synth.py seed 42)"
    else:
        g = odgi_b200.load_graph_arrays(os.path.join(ROOT, "tests", "golden", f"{name}.graph.arr.gz"))
        desc = f"test/{name}.gfa (flattened fixture)"
    return g, desc


def usable_cores():
    """host threads this process may actually use: the scheduler affinity mask, capped by the cgroup CPU quota.
    (os.cpu_count() is the machine's count: on a shared box it oversubscribes the reference's spinning workers.)"""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    quota = None
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:       # cgroup v2: "<quota|max> <period>"
            q, per = f.read().split()
            if q != "max":
                quota = float(q) / float(per)
    except (OSError, ValueError):
        try:
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f, open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as g:
                q, per = int(f.read()), int(g.read())
                if q > 0:
                    quota = q / per
        except (OSError, ValueError):
            pass
    if quota is not None:
        n = max(1, min(n, int(quota)))
    return n


def cpu_threads():
    """(threads to use, usable cores, 1-minute load average).  Inside a cpuset smaller than the machine (a 1-GPU lease: 16 of
    128 cores) the usable cores are ours — the machine-wide load average is other tenants on other cores.  On the whole
    machine the cores other processes keep busy are left to them: the reference's workers spin on per-node locks
    (src/odgi.cpp:65-71) and collapse when oversubscribed (round 1: 128 workers on a 16-core lease -> 9x slower)."""
    usable = usable_cores()
    machine = os.cpu_count() or usable
    try:
        load1 = os.getloadavg()[0]
    except OSError:
        load1 = 0.0
    if usable < machine:
        return usable, usable, load1
    free = usable - int(load1 + 0.5)
    return max(min(usable, 4), min(usable, free)), usable, load1


def cpu_baseline(workload: str, threads: int, iters: int = 2, warmup_iters: int = 0, single_thread_iters: int = 0):
    """Reference CPU implementation on the host cores, on a bounded sample of the workload.
    kind 'reference' = the unmodified reference compiled into oracle/_ref (std::thread workers, -Ofast);
    falls back to kind 'port' (the oracle's C restatement, one core) when oracle/_ref was not built.
    single_thread_iters > 0 adds a T = 1 run (T in {1, usable})."""
    from odgi_b200 import synth
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver_fast")
    if not os.path.exists(ref):
        ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if workload in synth.PRESETS:
        n_sites, n_paths = synth.PRESETS[workload]
        n_sites = min(n_sites, 30_000)  # bounded sample: same generator, same haplotype count, fewer sites
        g = synth.generate(n_sites, n_paths, seed=42)
        sample = f"same generator, {n_paths} paths x {n_sites} sites (S={g.S}), {iters} iterations of 10*S updates"
    else:
        import odgi_b200
        g = odgi_b200.load_graph_arrays(os.path.join(ROOT, "tests", "golden", f"{workload}.graph.arr.gz"))
        sample = f"the whole graph (S={g.S}), {iters} iterations of 10*S updates"
    if os.path.exists(ref):
        with tempfile.TemporaryDirectory() as tmp:
            gfa = os.path.join(tmp, "sample.gfa")
            synth.write_gfa(g, gfa)
            out = os.path.join(tmp, "o.arr")

            def run(t, n_iter, limit):
                r = subprocess.run([ref, "layout", gfa, "-", out, f"threads={t}", f"iter_max={n_iter}"], cwd=tmp, capture_output=True,
                                   text=True, timeout=limit)
                return json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else None

            if warmup_iters:  # untimed: page cache, CPU clocks
                run(threads, max(2, warmup_iters), 600)
            info = run(threads, iters, 1200)
            if info is not None:
                res = {"value": info["updates_per_sec"] / 1e6, "unit": "M updates/s", "cores": threads, "kind": "reference",
                       "sample": sample + f"; {os.path.basename(ref)} (reference sources, -Ofast)", "seconds": info["seconds"]}
                if single_thread_iters:
                    one = run(1, max(2, single_thread_iters), 1200)
                    if one is not None:
                        res["single_thread"] = {"value": one["updates_per_sec"] / 1e6, "cores": 1, "iterations": max(2, single_thread_iters),
                                                "seconds": one["seconds"]}
                return res
    from oracle import oracle as orc
    go = orc.Graph(g.node_len, g.path_first_step, g.step_node, g.step_rev)
    cfg = orc.default_layout_config(go, iter_max=iters)
    X, Y = orc.layout_init(go, 42)
    t0 = time.time()
    n, _, _ = orc.layout_2d(go, cfg, X, Y, n_streams=1)
    dt = time.time() - t0
    return {"value": n / dt / 1e6, "unit": "M updates/s", "cores": 1, "kind": "port", "sample": sample + "; oracle C restatement", "seconds": dt}


def reference_cuda_kernel(workload: str = "mid", iter_max: int = 30):
    """The reference's OWN CUDA kernel (src/cuda/layout.cu compiled unmodified for sm_100a: oracle/_ref/ref_gpu_driver),
    measured live on this GPU.  north_star names it as the reported baseline the >= 10x target is judged against.  Default
    graph 'mid' (same generator, 6e5 nodes / 4.6e7 steps): the driver needs the graph as GFA through odgi's own ingest,
    seconds there, minutes at c4 (recorded c4 measurement: profiles/r01_reference_cuda_kernel_c4.json)."""
    from odgi_b200 import synth
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_driver")
    if not os.path.exists(drv):
        return {"unavailable": "oracle/_ref/ref_gpu_driver was not built"}
    g = synth.preset(workload, seed=42)
    threads = max(4, cpu_threads()[0])
    with tempfile.TemporaryDirectory() as tmp:
        gfa = os.path.join(tmp, "g.gfa")
        t0 = time.time()
        synth.write_gfa(g, gfa)
        t_gfa = time.time() - t0
        r = subprocess.run([drv, gfa, "-", str(iter_max), str(threads)], cwd=tmp, capture_output=True, text=True, timeout=1500)
        if r.returncode != 0:
            return {"unavailable": f"ref_gpu_driver rc={r.returncode}: {r.stderr[-300:]}"}
        info = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    return {"value": info["updates_per_sec"] / 1e6, "unit": "M updates/s", "n_gpus": 1, "workload": workload, "nodes": info["nodes"],
            "steps": info["steps"], "iter_max": iter_max, "loop_s": info["loop_s"], "gfa_write_s": t_gfa, "gfa_load_s": info["gfa_load_s"],
            "how": "cuda::gpu_layout called with iter_max and 2*iter_max, loop time by difference (the reference has no hook around its loop)"}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads, usable, load1 = cpu_threads()
    # each step = one iteration (10*S updates) over a bounded sample of the workload (same generator, same haplotype count,
    # fewer sites): W untimed iterations, then K timed ones; the reference times its own SGD loop (graph load excluded)
    t0 = time.time()
    iters = max(2, args.steps)   # the reference's schedule divides by iter_max - 1
    cb = cpu_baseline(args.workload, threads, iters=iters, warmup_iters=max(0, min(args.warmup, 3)), single_thread_iters=2)
    cbl = {"value": cb["value"], "unit": cb["unit"], "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"],
           "usable_cores": usable, "machine_cores": os.cpu_count(), "loadavg_1m_before": load1}
    if "single_thread" in cb:
        cbl["single_thread"] = cb["single_thread"]
    line = {"impl": "reference", "metric": "M node-pair SGD updates/sec", "value": cb["value"], "unit": "M updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds"] / iters * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "sample": cb["sample"]},
            "cpu_baseline": cbl,
            "e2e": {"value": cb["value"], "unit": "M updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)


def run_reference_cuda_arm(args):
    """--impl reference-cuda: the reference's own CUDA kernel on this GPU (rank 0 only; it is a single-GPU program)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = args.workload if args.workload != "c4" or args.force_c4 else "mid"
    rk = reference_cuda_kernel(wl)
    line = {"impl": "reference-cuda", "metric": "M node-pair SGD updates/sec", "unit": "M updates/s", "n_gpus": 1, "higher_is_better": True,
            "dtype": "f32 storage, f64 arithmetic", "data": "synthetic", "config": {"workload": wl}}
    line.update(rk)
    print(json.dumps(line), flush=True)


def main():
    # the contract is ONE JSON line on stdout: anything libraries print there (NCCL's version banner, ...) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4", help="c4 (default) | mid | small | tiny | test-graph name | c5 | c5s (rank-locally generated, path-sharded: "
                                                    "BASELINE config 5 needs --gpus 8)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--force-c4", action="store_true", help="--impl reference-cuda on c4 itself (writes a 9 GB GFA, ~6 minutes)")
    ap.add_argument("--no-reference-cuda", action="store_true", help="skip the live reference-CUDA-kernel leg of the default line")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--multi", default="auto", choices=["auto", "hybrid", "peer", "allreduce", "single", "sharded"],
                    help="N>1: 'auto' (default) = allreduce where every replica still sees >= 60 updates per node and iteration (graphs many "
                         "haplotypes deep, c4), else single (rank 0 alone + one broadcast: shallow graphs leave the reference band in every shared mode); 'allreduce' = replicated coordinates + 1 NCCL all-reduce/step (north_star's design); "
                         "'peer' = coordinates partitioned over the GPUs, updated through NVLink peer memory (one shared Hogwild); 'hybrid' = allreduce "
                         "for the first third of the schedule, peer afterwards; 'sharded' = allreduce with the step records dealt out over the ranks by "
                         "path (capacity mode for graphs whose records do not fit one GPU; quality readout covers rank 0's paths)")
    ap.add_argument("--sampling", type=int, default=0, help="0 auto, 1 stream (reference-exact worker streams), 2 tile")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.impl == "reference-cuda":
        return run_reference_cuda_arm(args)

    import odgi_b200
    from odgi_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and not (world == 1 and args.gpus == 1):
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    import torch
    if not torch.cuda.is_available() or odgi_b200.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback on the hot path)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W, K = args.warmup, args.steps
    iter_max = max(30, W + K)
    from odgi_b200 import synth
    generated_sharded = args.workload in synth.SHARDED_PRESETS
    if generated_sharded:
        # BASELINE config 5 (and its small stand-in): every rank synthesises ITS OWN paths (p = rank mod world) plus the common node
        # table — no process ever holds the whole graph; step records are path-sharded, coordinates replicated, one all-reduce/step.
        # The reference's own CUDA kernel cannot run this size at all: 32-bit step index (src/cuda/layout.cu:207).
        n_sites, n_paths = synth.SHARDED_PRESETS[args.workload]
        g, my_paths = synth.generate_sharded(n_sites, n_paths, rank, world, seed=42)
        desc = f"synthetic '{args.workload}' generated rank-locally (odgi_b200/synth.py generate_sharded, seed 42): {n_sites} sites x {n_paths} haplotypes"
        tot = torch.tensor([float(g.S), float(g.max_path_steps)], dtype=torch.float64, device="cuda")
        if dist is not None:
            tmax = tot.clone()
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tot[1] = tmax[1]
        job_steps, job_max_steps = int(tot[0].item()), int(tot[1].item())
        # layout_main.cpp:251-266 on the WHOLE job: U = 10 * S, space = longest path in steps, eta_max = that squared
        cfg = capi.layout_defaults(g, iter_max=iter_max, batch=args.batch, n_streams=args.streams, flags=args.flags, sampling=args.sampling)
        cfg.min_term_updates = 10 * job_steps
        cfg.space = job_max_steps
        cfg.eta_max = float(job_max_steps) * float(job_max_steps)
        sharded = world > 1
        args.multi = "sharded" if sharded else args.multi
    else:
        g, desc = make_workload(args.workload)
        # the schedule's parameters (U, space, eta_max) are those of the WHOLE job, whatever a rank holds of it
        cfg = capi.layout_defaults(g, iter_max=iter_max, batch=args.batch, n_streams=args.streams, flags=args.flags, sampling=args.sampling)
        job_steps = g.S
        sharded = world > 1 and args.multi == "sharded"
        if sharded:
            g = odgi_b200.shard_paths(g, world, rank)   # this rank's paths only; node table whole
    multi_mode = {"auto": capi.MULTI_AUTO, "peer": capi.MULTI_PEER, "hybrid": capi.MULTI_HYBRID, "allreduce": capi.MULTI_ALLREDUCE,
                  "single": capi.MULTI_SINGLE, "sharded": capi.MULTI_ALLREDUCE}[args.multi]
    sampling_name = {1: "stream", 2: "tile"}.get(args.sampling, "tile" if g.S >= (1 << 22) else "stream")
    X0, Y0 = odgi_b200.layout_init(g, seed=42)
    U = cfg.min_term_updates

    uid = None
    if world > 1:
        obj = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        uid = obj[0]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident measurement ("value") ----
    e = odgi_b200.Engine(g, device=local_rank)
    if uid is not None:
        e.attach_comm(uid, world, rank)
        e.set_multi_mode(multi_mode)
        if sharded:
            e.set_shard(job_steps)
    e.set_coords_2d(X0, Y0)
    stress_initial = e.path_stress(2, 4_000_000, 12345)
    barrier()
    e.run_range(cfg, 2, 0, W)                      # W untimed warm-up steps (iterations 0..W-1 of the schedule)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.time()
    st = e.run_range(cfg, 2, W, W + K)             # exactly K timed steps
    torch.cuda.synchronize()
    wall = time.time() - t0
    barrier()
    clocks = sampler.stop()
    dev_s = max_over_ranks(st["seconds_iterations"])   # CUDA events around the whole call (both phases of a hybrid run + the switch)
    wall_s = max_over_ranks(wall)
    assert st["iterations_run"] == K and st["kernel_launches"] in (K, 0)   # 0: a rank that only receives the broadcast (single mode)
    # counted term updates of the whole job (every rank reports its own share)
    total_updates = st["term_updates"]
    if dist is not None:
        tt = torch.tensor([total_updates], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        total_updates = float(tt.item())
    assert abs(total_updates - K * U) <= 1e-3 * K * U, (total_updates, K * U)
    # 1 GPU: device time of the K launches (CUDA events on the kernels' stream).  N GPUs: the slowest rank's HOST wall clock
    # between the barriers — collectives, phase switches and host-side waits included.
    timed_s = dev_s if world == 1 else wall_s
    value = total_updates / timed_s / 1e6
    resolved = {capi.MULTI_ALLREDUCE: "allreduce", capi.MULTI_PEER: "peer", capi.MULTI_HYBRID: "hybrid", capi.MULTI_SINGLE: "single"}.get(e.resolved_multi_mode(), "?") if world > 1 else None
    if sharded:
        resolved = "sharded"
    # layout quality of the COMPLETE schedule: finish the remaining iterations (untimed) and evaluate the sampled path
    # stress on the device (collective in multi-GPU runs)
    quality = None
    if W + K <= iter_max:
        if W + K < iter_max:
            e.run_range(cfg, 2, W + K, iter_max)
        quality = {"stress_initial": stress_initial, "stress_final": e.path_stress(2, 4_000_000, 12345), "local_stress_final": e.local_stress(2, 4_000_000, 12345),
                   "pairs": 4_000_000,
                   "definition": "sampled path stress, SURVEY.md 8d / pgsgd_engine_path_stress; complete default schedule of " + str(iter_max) + " iterations"}
    Xf, Yf = e.get_coords_2d()
    finite = bool(np.all(np.isfinite(Xf)) and np.all(np.isfinite(Yf)))
    dev_bytes = e.device_bytes
    e.close()

    # ---- end-to-end through the C-ABI with host buffers ("e2e") ----
    e2e = None
    if not args.no_e2e:
        keep = []
        def pin(a):
            v, t = pinned_like(a)
            keep.append(t)
            return v
        gp = capi.FlatGraph(pin(g.node_len), pin(g.path_first_step), pin(g.step_node), None if g.step_rev is None else pin(g.step_rev),
                            None if g.step_pos is None else pin(g.step_pos))
        Xp, Yp = pin(X0), pin(Y0)
        Xr, Yr = pin(np.zeros_like(X0)), pin(np.zeros_like(Y0))   # the caller's result buffers (the shim fills the caller's vectors)
        cfg_e = cfg
        barrier()
        t0 = time.time()
        e2 = odgi_b200.Engine(gp, device=local_rank)        # flatten-to-device upload
        t_create = time.time()
        if uid is not None:
            obj = [capi.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            e2.attach_comm(obj[0], world, rank)
            e2.set_multi_mode(multi_mode)
            if sharded:
                e2.set_shard(job_steps)
        t_comm = time.time()
        e2.set_coords_2d(Xp, Yp)                            # coordinate upload
        t_up = time.time()
        st2 = e2.run_range(cfg_e, 2, 0, K)                  # the same number of steps
        t_run = time.time()
        Xo, Yo = e2.get_coords_2d(out=(Xr, Yr))             # result download into the caller's buffers
        t_down = time.time()
        phases = {"engine_create_flatten_upload_s": t_create - t0, "comm_attach_s": t_comm - t_create, "coords_upload_s": t_up - t_comm,
                  "run_call_s": t_run - t_up, "of_which_device_iterations_s": st2["seconds_iterations"], "coords_download_s": t_down - t_run}
        t_e2e = max_over_ranks(time.time() - t0)
        h2d = st2["h2d_bytes"]
        e2.close()
        tot2 = float(st2["term_updates"])
        if dist is not None:
            t2 = torch.tensor([tot2], dtype=torch.float64, device="cuda")
            dist.all_reduce(t2, op=dist.ReduceOp.SUM)
            tot2 = float(t2.item())
        e2e = {"value": tot2 / t_e2e / 1e6, "unit": "M updates/s", "h2d_bytes_per_step": h2d / K,
               "d2h_bytes_per_step": 4 * g.N * 8 / K, "seconds": t_e2e, "steps_in_call": K, "phases_rank0": phases,
               "note": "one engine lifetime: graph flatten+upload from pinned host memory, coords up, K steps, coords down; host wall clock"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    step_s = timed_s / K
    achieved = (total_updates / K / world) * BYTES_PER_UPDATE_2D / step_s / 1e9   # per-GPU kernel: its share of the step's updates
    # dram__bytes_read + dram__bytes_write per launch of the dominant kernel from the committed ncu --set full capture of the
    # 1-GPU launch (profiles/traffic_per_launch.json); a rank of an N-GPU run launches a different share, so null there
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_per_launch.json")
    if os.path.exists(tp) and world == 1:
        with open(tp) as f:
            traffic = json.load(f).get(args.workload)
    cb = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            threads, usable, load1 = cpu_threads()
            cb = cpu_baseline(args.workload, threads, iters=10, single_thread_iters=2)   # ~10-30 s of host CPU work
            cb["usable_cores"], cb["machine_cores"], cb["loadavg_1m_before"] = usable, os.cpu_count(), load1
        except Exception as ex:  # the baseline is a reported number, never a reason to lose the bench line
            cb = {"value": None, "unit": "M updates/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
    line = {
        "metric": "M node-pair SGD updates/sec", "value": value, "unit": "M updates/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic" if args.workload in ("c4", "c4x", "mid", "small", "tiny", "c5", "c5s") else "reference test graph (flattened fixture)",
        "config": {"workload": args.workload, "description": desc, "nodes": g.N, "paths": g.P if not sharded else None, "steps_in_graph": job_steps,
                   "updates_per_step": U, "iter_max": iter_max, "timed_iterations": [W, W + K], "sampling": sampling_name, "batch": args.batch or "auto",
                   "l2_policy": "inputs larger than L2" if g.S * 16 > 126e6 else "L2-resident graph (plumbing config)",
                   "parallelism": ("1 GPU" if world == 1 else
                                   f"coords partitioned over {world} GPUs, updated through NVLink peer memory (one shared Hogwild), tiles owned by node range"
                                   if resolved == "peer" else
                                   f"hybrid over {world} GPUs: iterations < {iter_max // 3} replicated + 1 NCCL all-reduce/step, then coords partitioned and updated through NVLink peer memory"
                                   if resolved == "hybrid" else
                                   f"rank 0 runs the whole job, one broadcast of the coordinates to the other {world - 1} GPUs"
                                   if resolved == "single" else
                                   f"step records dealt out over {world} GPUs by path (rank 0 holds {g.P} paths, {g.S} steps), replicated coords, 1 NCCL all-reduce/step"
                                   if sharded else f"replicated coords, term updates split over {world} GPU(s), 1 NCCL all-reduce/step"),
                   "multi_mode": resolved, "multi_requested": args.multi if world > 1 else None,
                   "timed": "CUDA events on the kernels' stream" if world == 1 else "host wall clock between barriers, max over ranks",
                   "device_bytes": dev_bytes, "coords_finite": finite},
        "gpu_launches": K * world,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "bytes_per_update": BYTES_PER_UPDATE_2D,
                     "kernel": ("pgsgd_tile_kernel<2,BATCH,...> (legacy)" if args.flags & 32 else "pgsgd_tile2_kernel<2,TMA,TILE>") if sampling_name == "tile"
                     else "pgsgd_iter_kernel<2,BATCH,smem_paths>",
                     "kernel_ms": step_s * 1e3},
        "clocks": clocks, "wall_s_timed_region": wall_s, "device_event_s_timed_region": dev_s,
        "rank0_kernel_s": st.get("seconds_kernels"), "rank0_collective_s": st.get("seconds_collectives"),
    }
    # The reference's own CUDA kernel (src/cuda/layout.cu compiled unmodified for sm_100a): the reported baseline north_star's
    # ">= 10x" is judged against.  Measured LIVE on this GPU on the 'mid' graph of the same generator, next to our kernel on
    # the same graph; the c4 number is a recorded measurement (its GFA ingest through odgi takes more than a minute).
    if world == 1 and not args.no_reference_cuda:
        try:  # context only: never a reason to lose the bench line
            rk = reference_cuda_kernel("mid")
            if "value" in rk:
                gm, _ = make_workload("mid")
                cm = capi.layout_defaults(gm, iter_max=30, flags=args.flags, sampling=args.sampling)
                Xm, Ym = odgi_b200.layout_init(gm, seed=42)
                with odgi_b200.Engine(gm, device=local_rank) as em:
                    em.set_coords_2d(Xm, Ym)
                    em.run_range(cm, 2, 0, 3)
                    sm = em.run_range(cm, 2, 3, 30)
                ours = sm["term_updates"] / sm["seconds_iterations"] / 1e6
                rk["ours_same_graph"] = {"value": ours, "unit": "M updates/s", "iterations": [3, 30]}
                rk["ratio_ours_over_reference_kernel"] = ours / rk["value"]
            rp = os.path.join(ROOT, "profiles", "r02_reference_cuda_kernel_c4.json")
            if args.workload == "c4" and os.path.exists(rp):
                with open(rp) as f:
                    rec = json.loads([ln for ln in f.read().splitlines() if ln.startswith("{")][-1])
                rk["recorded_c4"] = {"value": rec["updates_per_sec"] / 1e6, "unit": "M updates/s", "source": os.path.relpath(rp, ROOT),
                                     "ratio_ours_over_reference_kernel": value / (rec["updates_per_sec"] / 1e6),
                                     "note": "recorded in round 2 on this pool (scripts/gpu_runs/r02_call18.sh: same binary, same graph, 30 iterations; round 1 "
                                             "recorded 8348 with 10); not re-run here: 4 GB of GFA and 72 s of odgi ingest"}
            line["reference_cuda_kernel"] = rk
        except Exception as ex:
            line["reference_cuda_kernel"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
    if quality:
        line["quality"] = quality
    if e2e:
        line["e2e"] = e2e
    if cb:
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "single_thread", "usable_cores", "machine_cores",
                                                   "loadavg_1m_before") if k in cb}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
