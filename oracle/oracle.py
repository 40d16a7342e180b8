"""ctypes front-end to oracle/pgsgd_oracle.c (the CPU restatement of the reference's PG-SGD).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  Parity status: pinned against the reference
itself (tests/test_oracle_pinned.py, scripts/pin_oracle.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpgsgd_oracle.so")
_SRC = os.path.join(_HERE, "pgsgd_oracle.c")
_HDR = os.path.join(_HERE, "pgsgd_oracle.h")


def build(force: bool = False) -> str:
    """Compile the C restatement (strict IEEE: no fast-math, no FMA contraction)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= max(os.path.getmtime(_SRC), os.path.getmtime(_HDR)):
        return _SO
    cmd = ["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


class _Graph(C.Structure):
    _fields_ = [("node_count", C.c_uint64), ("path_count", C.c_uint64), ("step_count", C.c_uint64),
                ("node_len", C.c_void_p), ("path_first_step", C.c_void_p), ("step_node", C.c_void_p),
                ("step_rev", C.c_void_p), ("step_pos", C.c_void_p), ("step_perm", C.c_void_p)]


class _Config(C.Structure):
    _fields_ = [("iter_max", C.c_uint64), ("iter_with_max_learning_rate", C.c_uint64), ("min_term_updates", C.c_uint64),
                ("delta", C.c_double), ("eps", C.c_double), ("eta_max", C.c_double), ("theta", C.c_double),
                ("space", C.c_uint64), ("space_max", C.c_uint64), ("space_quantization_step", C.c_uint64),
                ("cooling_start", C.c_double), ("seed", C.c_uint64)]


class _Rng(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4)]


TERM_DTYPE = np.dtype([("step_index", "<u8"), ("path", "<u8"), ("rank_a", "<u8"), ("rank_b", "<u8"),
                       ("node_a", "<u4"), ("node_b", "<u4"), ("rev_a", "u1"), ("rev_b", "u1"), ("flip_a", "u1"),
                       ("flip_b", "u1"), ("end_a", "u1"), ("end_b", "u1"), ("_pad0", "u1", (2,)),
                       ("pos_a", "<u8"), ("pos_b", "<u8"), ("zipf", "u1"), ("_pad1", "u1", (7,))])

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_rng_seed.argtypes = [C.POINTER(_Rng), C.c_uint64]
        L.orc_rng_next.argtypes = [C.POINTER(_Rng)]
        L.orc_rng_next.restype = C.c_uint64
        L.orc_uniform.argtypes = [C.POINTER(_Rng), C.c_uint64]
        L.orc_uniform.restype = C.c_uint64
        L.orc_canonical.argtypes = [C.POINTER(_Rng)]
        L.orc_canonical.restype = C.c_double
        L.orc_fast_precise_pow.argtypes = [C.c_double, C.c_double]
        L.orc_fast_precise_pow.restype = C.c_double
        L.orc_zeta.argtypes = [C.c_uint64, C.c_double]
        L.orc_zeta.restype = C.c_double
        L.orc_dirty_zipf.argtypes = [C.POINTER(_Rng), C.c_uint64, C.c_double, C.c_double]
        L.orc_dirty_zipf.restype = C.c_uint64
        L.orc_schedule.argtypes = [C.c_double, C.c_uint64, C.c_uint64, C.c_double, C.c_void_p]
        L.orc_zetas.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_void_p, C.c_uint64]
        L.orc_zetas.restype = C.c_uint64
        L.orc_sample_term.argtypes = [C.POINTER(_Graph), C.POINTER(_Config), C.c_void_p, C.c_int, C.c_int, C.c_double,
                                      C.POINTER(_Rng), C.c_void_p]
        L.orc_sample_term.restype = C.c_int
        for name in ("orc_layout_2d",):
            getattr(L, name).argtypes = [C.POINTER(_Graph), C.POINTER(_Config), C.c_uint64, C.c_void_p, C.c_void_p]
            getattr(L, name).restype = C.c_uint64
        L.orc_layout_2d_f32.argtypes = [C.POINTER(_Graph), C.POINTER(_Config), C.c_uint64, C.c_void_p]
        L.orc_layout_2d_f32.restype = C.c_uint64
        L.orc_sort_1d.argtypes = [C.POINTER(_Graph), C.POINTER(_Config), C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_sort_1d.restype = C.c_uint64
        L.orc_run_range.argtypes = [C.POINTER(_Graph), C.POINTER(_Config), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_run_range.restype = C.c_uint64
        L.orc_run_inflight.argtypes = [C.POINTER(_Graph), C.POINTER(_Config), C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_run_inflight.restype = C.c_uint64
        L.orc_run_tile_order.argtypes = [C.POINTER(_Graph), C.POINTER(_Config), C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_run_tile_order.restype = C.c_uint64
        L.orc_peer_stale_2d_f32.argtypes = [C.POINTER(_Graph), C.POINTER(_Config), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_peer_stale_2d_f32.restype = C.c_uint64
        L.orc_replay_single.argtypes = [C.POINTER(_Graph), C.POINTER(_Config), C.c_int, C.c_uint64, C.c_uint64, C.c_double,
                                        C.c_double, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_replay_single.restype = C.c_uint64
        L.orc_replay_single_frozen.argtypes = L.orc_replay_single.argtypes + [C.c_void_p]
        L.orc_replay_single_frozen.restype = C.c_uint64
        L.orc_path_stress_2d.argtypes = [C.POINTER(_Graph), C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]
        L.orc_path_stress_2d.restype = C.c_double
        L.orc_path_stress_1d.argtypes = [C.POINTER(_Graph), C.c_void_p, C.c_uint64, C.c_uint64]
        L.orc_path_stress_1d.restype = C.c_double
        L.orc_local_stress_2d.argtypes = [C.POINTER(_Graph), C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]
        L.orc_local_stress_2d.restype = C.c_double
        L.orc_local_stress_1d.argtypes = [C.POINTER(_Graph), C.c_void_p, C.c_uint64, C.c_uint64]
        L.orc_local_stress_1d.restype = C.c_double
        assert C.sizeof(_Config) == 96
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class Config:
    """Mirror of the reference's PG-SGD parameter list (path_sgd_layout.hpp:37-56)."""
    iter_max: int = 30
    iter_with_max_learning_rate: int = 0
    min_term_updates: int = 0
    delta: float = 0.0
    eps: float = 0.01
    eta_max: float = 0.0
    theta: float = 0.99
    space: int = 0
    space_max: int = 1000
    space_quantization_step: int = 100
    cooling_start: float = 0.5
    seed: int = 9399220

    def c(self) -> _Config:
        return _Config(self.iter_max, self.iter_with_max_learning_rate, self.min_term_updates, self.delta, self.eps,
                       self.eta_max, self.theta, self.space, self.space_max, self.space_quantization_step,
                       self.cooling_start, self.seed)


class Graph:
    """Path-major flattened graph held as numpy arrays (kept alive for the C side)."""

    def __init__(self, node_len, path_first_step, step_node, step_rev, step_pos=None, step_perm=None):
        self.node_len = np.ascontiguousarray(node_len, dtype=np.uint32)
        self.path_first_step = np.ascontiguousarray(path_first_step, dtype=np.uint64)
        self.step_node = np.ascontiguousarray(step_node, dtype=np.uint32)
        self.step_rev = np.ascontiguousarray(step_rev, dtype=np.uint8)
        if step_pos is None:
            step_pos = positions_from_lengths(self.node_len, self.path_first_step, self.step_node)
        self.step_pos = np.ascontiguousarray(step_pos, dtype=np.uint64)
        self.step_perm = None if step_perm is None else np.ascontiguousarray(step_perm, dtype=np.uint64)
        self.N = int(self.node_len.size)
        self.P = int(self.path_first_step.size - 1)
        self.S = int(self.step_node.size)

    @classmethod
    def from_arrays(cls, arrs, use_xp_perm: bool = False):
        perm = None
        if use_xp_perm:
            # XP's node-major sampling table: step_index -> (path id, 1-based rank) (xp.cpp:143-144)
            first = arrs["path_first_step"]
            pid = arrs["xp_npi_iv"].astype(np.int64)
            path_index = {int(v): i for i, v in enumerate(arrs["xp_path_id"])}
            pidx = np.array([path_index[int(v)] for v in pid], dtype=np.uint64)
            perm = first[pidx] + arrs["xp_nr_iv"] - 1
        return cls(arrs["node_len"], arrs["path_first_step"], arrs["step_node"], arrs["step_rev"], arrs["step_pos"], perm)

    def with_perm(self, perm):
        return Graph(self.node_len, self.path_first_step, self.step_node, self.step_rev, self.step_pos, perm)

    def c(self) -> _Graph:
        return _Graph(self.N, self.P, self.S, _ptr(self.node_len), _ptr(self.path_first_step), _ptr(self.step_node),
                      _ptr(self.step_rev), _ptr(self.step_pos), _ptr(self.step_perm))

    @property
    def step_counts(self):
        return np.diff(self.path_first_step.astype(np.int64))

    @property
    def max_path_steps(self) -> int:
        return int(self.step_counts.max()) if self.P else 0

    @property
    def max_path_bp(self) -> int:
        best = 0
        for p in range(self.P):
            a, b = int(self.path_first_step[p]), int(self.path_first_step[p + 1])
            if b > a:
                best = max(best, int(self.step_pos[b - 1]) + int(self.node_len[self.step_node[b - 1]]))
        return best


def positions_from_lengths(node_len, path_first_step, step_node) -> np.ndarray:
    """step_pos[s] = bp offset of step s's node start within its path (XPPath ctor, xp.cpp:607-616)."""
    lens = node_len[step_node].astype(np.uint64)
    csum = np.cumsum(lens, dtype=np.uint64)
    pos = np.empty_like(csum)
    pos[0:1] = 0
    pos[1:] = csum[:-1]
    first = np.asarray(path_first_step[:-1], dtype=np.int64)
    counts = np.diff(np.asarray(path_first_step, dtype=np.int64))
    nonempty = counts > 0
    base = np.zeros(len(first), dtype=np.uint64)
    base[nonempty] = pos[first[nonempty]]
    return pos - np.repeat(base, counts)


def default_layout_config(g: Graph, **kw) -> Config:
    """`odgi layout` defaults (layout_main.cpp:198-266)."""
    ms = g.max_path_steps
    c = Config(iter_max=30, min_term_updates=10 * g.S, eta_max=float(ms) * float(ms), space=ms, space_max=1000,
               space_quantization_step=100)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def default_sort_config(g: Graph, **kw) -> Config:
    """`odgi sort -Y` defaults (sort_main.cpp:313-414)."""
    ms = g.max_path_steps
    space = g.max_path_bp
    space_max = 100
    max_dists = max(space_max + 1, 100)
    q = max(2, int(np.ceil((space - space_max) / (max_dists - space_max)))) if space > space_max and max_dists > space_max else 100   # sort_main.cpp:402-411
    c = Config(iter_max=100, min_term_updates=g.S, eta_max=float(ms) * float(ms), space=space, space_max=space_max,
               space_quantization_step=q)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def schedule(cfg: Config) -> np.ndarray:
    etas = np.zeros(cfg.iter_max + 1, dtype=np.float64)
    lib().orc_schedule(cfg.eta_max, cfg.iter_max, cfg.iter_with_max_learning_rate, cfg.eps, _ptr(etas))
    return etas


def zetas(cfg: Config) -> np.ndarray:
    n = lib().orc_zetas(cfg.space, cfg.space_max, cfg.space_quantization_step, cfg.theta, None, 0)
    z = np.zeros(n, dtype=np.float64)
    lib().orc_zetas(cfg.space, cfg.space_max, cfg.space_quantization_step, cfg.theta, _ptr(z), n)
    return z


def layout_init(g: Graph, seed: int = 42, noise: bool = True):
    """'d' initialisation (layout_main.cpp:322-328): X = cumulative bp, Y ~ N(0, sqrt(2N)).
    The reference seeds its mt19937 from std::random_device; parity runs inject this seeded one."""
    N = g.N
    X = np.zeros(2 * N, dtype=np.float64)
    csum = np.cumsum(g.node_len.astype(np.uint64))
    X[1::2] = csum
    X[2::2] = csum[:-1]
    if noise:
        rng = np.random.Generator(np.random.MT19937(seed))
        Y = rng.normal(0.0, np.sqrt(2.0 * N), size=2 * N)
    else:
        Y = np.zeros(2 * N, dtype=np.float64)
    return X, Y


def sort_init(g: Graph) -> np.ndarray:
    """1D initialisation: X[rank] = cumulative bp of the node order (path_sgd.cpp:63-69)."""
    csum = np.cumsum(g.node_len.astype(np.uint64))
    X = np.zeros(g.N, dtype=np.float64)
    X[1:] = csum[:-1]
    return X


def layout_2d(g: Graph, cfg: Config, X, Y, n_streams: int = 1) -> int:
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    gc, cc = g.c(), cfg.c()
    n = lib().orc_layout_2d(C.byref(gc), C.byref(cc), n_streams, _ptr(X), _ptr(Y))
    return int(n), X, Y


def layout_2d_f32(g: Graph, cfg: Config, xy, n_streams: int = 1):
    xy = np.ascontiguousarray(xy, dtype=np.float32)
    gc, cc = g.c(), cfg.c()
    n = lib().orc_layout_2d_f32(C.byref(gc), C.byref(cc), n_streams, _ptr(xy))
    return int(n), xy


def sort_1d(g: Graph, cfg: Config, X, n_streams: int = 1, frozen=None):
    X = np.ascontiguousarray(X, dtype=np.float64)
    fz = None if frozen is None else np.ascontiguousarray(frozen, dtype=np.uint8)
    gc, cc = g.c(), cfg.c()
    n = lib().orc_sort_1d(C.byref(gc), C.byref(cc), n_streams, _ptr(fz), _ptr(X))
    return int(n), X


def run_range(g: Graph, cfg: Config, n_streams: int, seed_base: int, updates: int, iter_begin: int, iter_end: int, mode: int,
              X=None, Y=None, xy=None, frozen=None, rng_state=None) -> int:
    """iterations [iter_begin, iter_end) in place on the given (contiguous, correctly typed) coordinate arrays"""
    gc, cc = g.c(), cfg.c()
    return int(lib().orc_run_range(C.byref(gc), C.byref(cc), n_streams, seed_base, updates, iter_begin, iter_end, mode,
                                   _ptr(X), _ptr(Y), _ptr(xy), _ptr(frozen), _ptr(rng_state)))


def run_tile_order(g: Graph, cfg: Config, tile_steps: int, dims: int, xy=None, X=None) -> int:
    """Sequential tile-ORDER model of the device's tile sampling (orc_run_tile_order), exact partner law.  In place."""
    gc, cc = g.c(), cfg.c()
    return int(lib().orc_run_tile_order(C.byref(gc), C.byref(cc), tile_steps, 1 if dims == 2 else 2, _ptr(xy), _ptr(X)))


def run_inflight(g: Graph, cfg: Config, n_streams: int, dims: int, exch: bool, xy=None, X=None) -> int:
    """Hogwild staleness model (orc_run_inflight): waves of n_streams terms that all read before any of them writes;
    exch=False sums the displacements (red.add), True is last-writer-wins.  In place on xy (2D fp32) or X (1D)."""
    gc, cc = g.c(), cfg.c()
    return int(lib().orc_run_inflight(C.byref(gc), C.byref(cc), n_streams, 1 if dims == 2 else 2, 1 if exch else 0, _ptr(xy), _ptr(X)))


def emulate_sharded_2d_f32(shards, global_steps: int, cfg: Config, xy0, n_streams: int):
    """Single-process emulation of the PATH-SHARDED multi-GPU schedule (pgsgd_engine_set_shard): rank r holds only the
    graph shards[r] (its paths), performs U * S_r / S updates per iteration on it with worker streams
    seed + r*n_streams + t, then the replicas are averaged in fp32."""
    n_ranks = len(shards)
    reps = [np.ascontiguousarray(xy0, dtype=np.float32).copy() for _ in range(n_ranks)]
    states = [np.zeros(4 * n_streams, dtype=np.uint64) for _ in range(n_ranks)]
    for it in range(cfg.iter_max):
        for r in range(n_ranks):
            share = cfg.min_term_updates * shards[r].S // global_steps
            run_range(shards[r], cfg, n_streams, cfg.seed + r * n_streams, share, it, it + 1, 1, xy=reps[r], rng_state=states[r])
        acc = reps[0].copy()
        for r in range(1, n_ranks):
            acc = acc + reps[r]
        merged = (acc / np.float32(n_ranks)).astype(np.float32)
        for r in range(n_ranks):
            reps[r][...] = merged
    return reps[0]


def emulate_multirank_2d_f32(g: Graph, cfg: Config, xy0, n_ranks: int, n_streams: int, sum_deltas: bool = False,
                             syncs_per_iter: int = 1):
    """Single-process emulation of the multi-GPU schedule of pgsgd_engine_run_2d with a communicator attached:
    every iteration is cut into syncs_per_iter slices; in each slice every rank performs its share of the slice's
    updates with its own worker streams (seed + rank*n_streams + t) on its replica, then the replicas are averaged in
    fp32 (all-reduce AVG) or the displacements are summed.  Returns the common coordinates after the last iteration."""
    reps = [np.ascontiguousarray(xy0, dtype=np.float32).copy() for _ in range(n_ranks)]
    states = [np.zeros(4 * n_streams, dtype=np.uint64) for _ in range(n_ranks)]
    U = cfg.min_term_updates
    for it in range(cfg.iter_max):
        for k in range(syncs_per_iter):
            u_slice = U // syncs_per_iter + (1 if k < U % syncs_per_iter else 0)
            prev = reps[0].copy()
            for r in range(n_ranks):
                share = u_slice // n_ranks + (1 if r < u_slice % n_ranks else 0)
                run_range(g, cfg, n_streams, cfg.seed + r * n_streams, share, it, it + 1, 1, xy=reps[r], rng_state=states[r])
            if sum_deltas:
                acc = np.zeros_like(prev)
                for r in range(n_ranks):
                    acc = acc + (reps[r] - prev)
                merged = prev + acc
            else:
                acc = reps[0].copy()
                for r in range(1, n_ranks):
                    acc = acc + reps[r]
                merged = (acc / np.float32(n_ranks)).astype(np.float32)
            for r in range(n_ranks):
                reps[r][...] = merged
    return reps[0]


def peer_stale_2d_f32(g: Graph, cfg: Config, xy, n_ranks: int, streams_per_rank: int, refreshes: int, iter_begin: int, iter_end: int) -> int:
    gc, cc = g.c(), cfg.c()
    return int(lib().orc_peer_stale_2d_f32(C.byref(gc), C.byref(cc), n_ranks, streams_per_rank, refreshes, iter_begin, iter_end, _ptr(xy)))


def replay_single(g: Graph, cfg: Config, dims: int, n_terms: int, switch_at: int, eta0: float, eta1: float,
                  cooling0: bool, cooling1: bool, theta1: float, X=None, Y=None, want_terms: bool = True, frozen=None):
    terms = np.zeros(n_terms, dtype=TERM_DTYPE) if want_terms else None
    fz = None if frozen is None else np.ascontiguousarray(frozen, dtype=np.uint8)
    gc, cc = g.c(), cfg.c()
    lib().orc_replay_single_frozen(C.byref(gc), C.byref(cc), dims, n_terms, switch_at, eta0, eta1, int(cooling0), int(cooling1),
                                   theta1, _ptr(X), _ptr(Y), _ptr(terms), _ptr(fz))
    return terms


def sample_terms(g: Graph, cfg: Config, dims: int, cooling: bool, n_terms: int, stream: int = 0, theta_zipf=None):
    """The first n_terms draws of worker stream `stream` (seed + stream), including the 1-step-path
    skips as rows with valid == 0."""
    L = lib()
    z = zetas(cfg)
    rng = _Rng()
    L.orc_rng_seed(C.byref(rng), cfg.seed + stream)
    terms = np.zeros(n_terms, dtype=TERM_DTYPE)
    valid = np.zeros(n_terms, dtype=np.uint8)
    gc, cc = g.c(), cfg.c()
    th = cfg.theta if theta_zipf is None else theta_zipf
    base = terms.ctypes.data
    for k in range(n_terms):
        valid[k] = L.orc_sample_term(C.byref(gc), C.byref(cc), _ptr(z), dims, int(cooling), th, C.byref(rng),
                                     C.c_void_p(base + k * TERM_DTYPE.itemsize))
    return terms, valid


def path_stress_2d(g: Graph, X, Y, n_pairs: int = 1_000_000, seed: int = 12345) -> float:
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    gc = g.c()
    return float(lib().orc_path_stress_2d(C.byref(gc), _ptr(X), _ptr(Y), n_pairs, seed))


def path_stress_1d(g: Graph, X, n_pairs: int = 1_000_000, seed: int = 12345) -> float:
    X = np.ascontiguousarray(X, dtype=np.float64)
    gc = g.c()
    return float(lib().orc_path_stress_1d(C.byref(gc), _ptr(X), n_pairs, seed))


def local_stress_2d(g: Graph, X, Y, n_pairs: int = 1_000_000, seed: int = 12345) -> float:
    """near-pair stress: partner 1..64 ranks away along the path, d <= 1000 bp (orc_local_stress_2d)"""
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    gc = g.c()
    return float(lib().orc_local_stress_2d(C.byref(gc), _ptr(X), _ptr(Y), n_pairs, seed))


def local_stress_1d(g: Graph, X, n_pairs: int = 1_000_000, seed: int = 12345) -> float:
    X = np.ascontiguousarray(X, dtype=np.float64)
    gc = g.c()
    return float(lib().orc_local_stress_1d(C.byref(gc), _ptr(X), n_pairs, seed))


def xy_to_XY(xy: np.ndarray):
    """float4-per-node device layout {x0,y0,x1,y1} -> reference X/Y indexed 2*node+end."""
    a = np.asarray(xy).reshape(-1, 4)
    X = np.empty(2 * a.shape[0], dtype=np.float64)
    Y = np.empty_like(X)
    X[0::2], Y[0::2], X[1::2], Y[1::2] = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    return X, Y


def XY_to_xy(X, Y, dtype=np.float32) -> np.ndarray:
    N = len(X) // 2
    a = np.empty((N, 4), dtype=dtype)
    a[:, 0], a[:, 1], a[:, 2], a[:, 3] = X[0::2], Y[0::2], X[1::2], Y[1::2]
    return a.reshape(-1)


def component_keys(n_nodes: int, path_first_step, step_node, edges=()) -> np.ndarray:
    """The weak-component key path_linear_sgd_order sorts by (path_sgd.cpp:552-586): weakly connected components of the
    graph (consecutive path steps are edges; `edges` adds node-rank pairs the paths do not cover, i.e. the remaining L
    lines), numbered by ascending average node id.  Returns the key of every node rank."""
    parent = np.arange(n_nodes, dtype=np.int64)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    first = np.asarray(path_first_step, dtype=np.int64)
    sn = np.asarray(step_node, dtype=np.int64)
    pairs = [(int(a), int(b)) for a, b in edges]
    for p in range(len(first) - 1):
        seg = sn[first[p]:first[p + 1]]
        pairs.extend(zip(seg[:-1].tolist(), seg[1:].tolist()))
    for a, b in pairs:
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    root = np.array([find(i) for i in range(n_nodes)])
    roots = np.unique(root)
    avg = sorted((float(np.mean(np.nonzero(root == r)[0] + 1)), int(r)) for r in roots)   # node id = rank + 1
    key_of = {r: k for k, (_, r) in enumerate(avg)}
    return np.array([key_of[int(r)] for r in root], dtype=np.uint32)


def order_from_x(x: np.ndarray, component: Optional[np.ndarray] = None) -> np.ndarray:
    """path_linear_sgd_order's sort (path_sgd.cpp:650-658): by (weak component, pos, handle integer).
    The reference clear()s weak_components_map before reading it (path_sgd.cpp:588), but clear() leaves the storage in
    place and the unchecked reads still return the component ids: the running reference DOES sort by component
    (tests/golden/order_multi3.json, made by scripts/make_order_golden.py).  component = component_keys(...) or None
    for a single-component graph.  Returns node ranks in sorted order."""
    n = len(x)
    comp = np.zeros(n, dtype=np.uint64) if component is None else np.asarray(component, dtype=np.uint64)
    handle = np.arange(n, dtype=np.uint64) << np.uint64(1)
    return np.lexsort((handle, np.asarray(x, dtype=np.float64), comp)).astype(np.uint64)


def sort_goodness(node_len, path_first_step, step_node, step_rev, order=None, dont_penalize_gap_links=True, penalize_diff_orientation=True):
    """`odgi stats -l [-g] -s [-d]` on the graph sorted by `order` (node ranks in sorted order; None = the graph as it is):
    mean_links_length and sum_of_path_node_distances in node space and nucleotide space, restated from
    src/subcommand/stats_main.cpp:399-800 (1D branch; ids compact, so nspace[k] == k).  The reference publishes these for
    DRB1-3123 before and after `odgi sort -Y` (docs/rst/tutorials/sort_layout.rst:101-106,175-186)."""
    node_len = np.asarray(node_len, dtype=np.int64)
    first = np.asarray(path_first_step, dtype=np.int64)
    sn = np.asarray(step_node, dtype=np.int64)
    rev = np.zeros(len(sn), dtype=np.int64) if step_rev is None else np.asarray(step_rev, dtype=np.int64)
    n = len(node_len)
    order = np.arange(n, dtype=np.int64) if order is None else np.asarray(order, dtype=np.int64)
    new_rank = np.empty(n, dtype=np.int64)
    new_rank[order] = np.arange(n, dtype=np.int64)
    pm = np.concatenate(([0], np.cumsum(node_len[order])))          # position_map (:429-447): bp before sorted rank k
    r = new_rank[sn]
    S = len(sn)
    is_last = np.zeros(S, dtype=bool)
    is_last[first[1:][first[1:] > first[:-1]] - 1] = True
    src = np.nonzero(~is_last)[0]                                   # steps that have a next step in their path
    uh, ui, bh, bi = r[src], r[src + 1], rev[src], rev[src + 1]
    # ---- mean_links_length (:449-600) ----
    gap = np.zeros(len(src), dtype=bool)
    if dont_penalize_gap_links:
        path_of = np.searchsorted(first, src, side="right") - 1
        for p in range(len(first) - 1):
            sel = path_of == p
            if not sel.any():
                continue
            uniq = np.unique(r[first[p]:first[p + 1]])             # the ordered set of node numbers in the path (:479,487-491)
            idx = np.searchsorted(uniq, uh[sel])
            succ = np.where(idx + 1 < len(uniq), uniq[np.minimum(idx + 1, len(uniq) - 1)], -1)
            gap[sel] = succ == ui[sel]
    ia = uh + (1 - bh)
    ib = ui + bi
    lo, hi = np.minimum(ia, ib), np.maximum(ia, ib)
    pen = ~gap
    links = len(src)
    mll_node = float((hi - lo)[pen].sum()) / links if links else 0.0
    mll_nt = float((pm[hi] - pm[lo])[pen].sum()) / links if links else 0.0
    # ---- sum_of_path_node_distances (:602-800) ----
    a, b = np.minimum(uh, ui), np.maximum(uh, ui)
    back = ui < uh
    w = np.where(back, 3, 1)
    d_node, d_nt = b - a, pm[b] - pm[a]
    sum_node = int((w * d_node).sum())
    sum_nt = int((w * d_nt).sum())
    diff = bh != bi
    if penalize_diff_orientation:
        sum_node += int(2 * d_node[diff].sum())
        sum_nt += int(2 * d_nt[diff].sum())
    nonempty = first[1:] > first[:-1]
    last_nodes = sn[first[1:][nonempty] - 1]
    sum_node += int(nonempty.sum())                                 # "add end of path so the best metric equals 1" (:733-735)
    sum_nt += int(node_len[last_nodes].sum())
    len_node, len_nt = S, int(node_len[sn].sum())
    return {"mean_links_length_node": mll_node, "mean_links_length_nt": mll_nt, "num_links": links, "num_gap_links": int(gap.sum()),
            "sum_path_node_dist_node": sum_node / len_node if len_node else 0.0, "sum_path_node_dist_nt": sum_nt / len_nt if len_nt else 0.0,
            "nodes": len_node, "nucleotides": len_nt, "num_penalties": int(back.sum()), "num_penalties_diff_orientation": int(diff.sum())}
