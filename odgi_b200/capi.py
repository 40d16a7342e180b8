"""ctypes binding of the C-ABI (include/pgsgd.h) — the same binding a C/C++ host makes.

There is deliberately no fallback here: if libpgsgd_b200.so is missing or no CUDA device is usable the
calls raise.  Nothing in this package imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpgsgd_b200.so")

PGSGD_FLAG_EXCH_WRITE = 1
PGSGD_FLAG_SUM_DELTAS = 2
PGSGD_FLAG_PLAIN_STORE = 4
PGSGD_FLAG_TMA_STAGING = 8
PGSGD_FLAG_KEEP_ADD = 16
SAMPLING_AUTO, SAMPLING_STREAM, SAMPLING_TILE = 0, 1, 2
FLAG_EXCH_WRITE, FLAG_SUM_DELTAS, FLAG_PLAIN_STORE, FLAG_TMA_STAGING, FLAG_KEEP_ADD, FLAG_LEGACY_TILE, FLAG_HALF_TILE, FLAG_BIG_TILE, FLAG_SWEEP_TILES, FLAG_L2_WINDOW = 1, 2, 4, 8, 16, 32, 64, 128, 256, 512
FLAG_WINDOW_TILES, FLAG_X_TILE_REPLACE, FLAG_X_STEP_RANDOM, FLAG_X_SEGMENT_RANDOM, FLAG_X_STEP_SCRAMBLE = 4096, 8192, 16384, 32768, 65536
FLAG_X_SCRAMBLE_PAIRS, FLAG_X_SCRAMBLE_QUADS = 131072, 262144   # with FLAG_X_STEP_SCRAMBLE: groups of 2 / 4 neighbouring lanes stay together
MULTI_ALLREDUCE, MULTI_PEER, MULTI_HYBRID, MULTI_AUTO, MULTI_SINGLE = 0, 1, 2, 3, 4


class PgsgdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pgsgd error {code}: {msg}")
        self.code = code


class GraphView(C.Structure):
    _fields_ = [("node_count", C.c_uint64), ("path_count", C.c_uint64), ("step_count", C.c_uint64),
                ("node_len", C.c_void_p), ("path_first_step", C.c_void_p), ("step_node", C.c_void_p),
                ("step_rev", C.c_void_p), ("step_pos", C.c_void_p)]


class ConfigC(C.Structure):
    _fields_ = [("iter_max", C.c_uint64), ("iter_with_max_learning_rate", C.c_uint64), ("min_term_updates", C.c_uint64),
                ("delta", C.c_double), ("eps", C.c_double), ("eta_max", C.c_double), ("theta", C.c_double),
                ("space", C.c_uint64), ("space_max", C.c_uint64), ("space_quantization_step", C.c_uint64),
                ("cooling_start", C.c_double), ("seed", C.c_uint64), ("n_streams", C.c_uint32), ("batch", C.c_uint32),
                ("flags", C.c_uint32), ("sampling", C.c_uint32), ("multi_switch_iteration", C.c_uint64)]


class StatsC(C.Structure):
    _fields_ = [("iterations_run", C.c_uint64), ("term_updates", C.c_uint64), ("seconds_iterations", C.c_double),
                ("seconds_upload", C.c_double), ("seconds_download", C.c_double), ("last_delta_max", C.c_double),
                ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("flags_used", C.c_uint64),
                ("sampling_used", C.c_uint64), ("seconds_kernels", C.c_double), ("seconds_collectives", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/pgsgd.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "pgsgd_last_error", "pgsgd_version", "pgsgd_device_count", "pgsgd_device_warmup", "pgsgd_layout_2d", "pgsgd_sort_1d", "pgsgd_layout_2d_multi",
    "pgsgd_sort_1d_multi",
    "pgsgd_engine_create", "pgsgd_engine_create_from_gfa_paths", "pgsgd_engine_graph_stats", "pgsgd_engine_destroy", "pgsgd_engine_device", "pgsgd_engine_device_bytes",
    "pgsgd_engine_set_coords_2d", "pgsgd_engine_get_coords_2d", "pgsgd_engine_set_coords_2d_f32",
    "pgsgd_engine_get_coords_2d_f32", "pgsgd_engine_set_coords_1d", "pgsgd_engine_get_coords_1d",
    "pgsgd_engine_set_frozen_1d", "pgsgd_engine_run_2d", "pgsgd_engine_run_1d", "pgsgd_engine_run_range", "pgsgd_comm_unique_id",
    "pgsgd_engine_attach_comm", "pgsgd_engine_set_multi_mode", "pgsgd_engine_resolved_multi_mode", "pgsgd_engine_set_shard", "pgsgd_engine_path_stress", "pgsgd_engine_local_stress", "pgsgd_engine_order_1d", "pgsgd_engine_order_1d_components", "pgsgd_engine_sort_goodness", "pgsgd_engine_encode_lay", "pgsgd_engine_sample_terms", "pgsgd_engine_set_trace", "pgsgd_engine_get_trace", "pgsgd_schedule", "pgsgd_zetas",
]

class GoodnessC(C.Structure):
    _fields_ = [("mean_links_length_node", C.c_double), ("mean_links_length_nt", C.c_double), ("num_links", C.c_uint64),
                ("num_gap_links", C.c_uint64), ("sum_path_node_dist_node", C.c_double), ("sum_path_node_dist_nt", C.c_double),
                ("nodes", C.c_uint64), ("nucleotides", C.c_uint64), ("num_penalties", C.c_uint64),
                ("num_penalties_diff_orientation", C.c_uint64)]


ABI_VERSION = 105  # PGSGD_VERSION of the include/pgsgd.h these ctypes structs mirror
_lib = None


def _prefer_environment_nccl():
    """libpgsgd_b200.so depends on libnccl.so.2 by soname, and a process holds ONE library per soname.  If this Python
    environment ships a (newer) NCCL next to PyTorch, load that one first: otherwise the system NCCL loaded for us would also
    be handed to a later `import torch`, which needs symbols of the version it was built against.  Plumbing only: the
    library itself runs with either."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
            return


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PgsgdError(-2, f"{LIB_PATH} is missing: build it with `python -m odgi_b200.build` (no CPU fallback exists)")
        _prefer_environment_nccl()
        L = C.CDLL(LIB_PATH)
        vp, u64, i32, dbl = C.c_void_p, C.c_uint64, C.c_int, C.c_double
        L.pgsgd_last_error.restype = C.c_char_p
        L.pgsgd_version.restype = i32
        if L.pgsgd_version() != ABI_VERSION:
            raise PgsgdError(-2, f"{LIB_PATH} has ABI {L.pgsgd_version()}, this binding mirrors include/pgsgd.h version {ABI_VERSION}: "
                                 "rebuild with `python -m odgi_b200.build`")
        L.pgsgd_device_count.restype = i32
        L.pgsgd_device_warmup.argtypes = [i32]
        L.pgsgd_layout_2d.argtypes = [C.POINTER(GraphView), C.POINTER(ConfigC), vp, vp, C.POINTER(StatsC)]
        L.pgsgd_sort_1d.argtypes = [C.POINTER(GraphView), C.POINTER(ConfigC), vp, i32, vp, C.POINTER(StatsC)]
        L.pgsgd_layout_2d_multi.argtypes = [C.POINTER(GraphView), C.POINTER(ConfigC), i32, i32, vp, vp, C.POINTER(StatsC)]
        L.pgsgd_sort_1d_multi.argtypes = [C.POINTER(GraphView), C.POINTER(ConfigC), i32, i32, vp, i32, vp, C.POINTER(StatsC)]
        L.pgsgd_engine_create.argtypes = [C.POINTER(GraphView), i32, C.POINTER(vp)]
        L.pgsgd_engine_create_from_gfa_paths.argtypes = [vp, u64, vp, vp, vp, u64, i32, C.POINTER(vp)]
        L.pgsgd_engine_graph_stats.argtypes = [vp, vp, vp, vp, vp]
        L.pgsgd_engine_destroy.argtypes = [vp]
        L.pgsgd_engine_destroy.restype = None
        L.pgsgd_engine_device.argtypes = [vp]
        L.pgsgd_engine_device_bytes.argtypes = [vp]
        L.pgsgd_engine_device_bytes.restype = u64
        for n in ("pgsgd_engine_set_coords_2d", "pgsgd_engine_get_coords_2d"):
            getattr(L, n).argtypes = [vp, vp, vp]
        for n in ("pgsgd_engine_set_coords_2d_f32", "pgsgd_engine_get_coords_2d_f32", "pgsgd_engine_set_coords_1d",
                  "pgsgd_engine_get_coords_1d", "pgsgd_engine_set_frozen_1d"):
            getattr(L, n).argtypes = [vp, vp]
        for n in ("pgsgd_engine_run_2d", "pgsgd_engine_run_1d"):
            getattr(L, n).argtypes = [vp, C.POINTER(ConfigC), C.POINTER(StatsC)]
        L.pgsgd_engine_run_range.argtypes = [vp, C.POINTER(ConfigC), i32, u64, u64, C.POINTER(StatsC)]
        L.pgsgd_comm_unique_id.argtypes = [vp]
        L.pgsgd_engine_attach_comm.argtypes = [vp, vp, i32, i32]
        L.pgsgd_engine_set_multi_mode.argtypes = [vp, i32]
        L.pgsgd_engine_set_shard.argtypes = [vp, u64]
        L.pgsgd_engine_resolved_multi_mode.argtypes = [vp]
        L.pgsgd_engine_path_stress.argtypes = [vp, i32, u64, u64, vp]
        L.pgsgd_engine_order_1d.argtypes = [vp, vp]
        L.pgsgd_engine_order_1d_components.argtypes = [vp, vp, vp]
        L.pgsgd_engine_sort_goodness.argtypes = [vp, vp, C.c_uint32, vp]
        L.pgsgd_engine_encode_lay.argtypes = [vp, vp, C.c_uint32, vp, u64, vp]
        L.pgsgd_engine_local_stress.argtypes = [vp, i32, u64, u64, vp]
        L.pgsgd_engine_sample_terms.argtypes = [vp, C.POINTER(ConfigC), i32, i32, dbl, u64, u64] + [vp] * 11
        L.pgsgd_engine_set_trace.argtypes = [vp, u64]
        L.pgsgd_engine_get_trace.argtypes = [vp, vp, vp, vp, vp]
        L.pgsgd_schedule.argtypes = [C.POINTER(ConfigC), vp]
        L.pgsgd_zetas.argtypes = [C.POINTER(ConfigC), vp, u64]
        L.pgsgd_zetas.restype = u64
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise PgsgdError(rc, lib().pgsgd_last_error().decode())


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class Config:
    """Argument list of the reference's path_linear_sgd_layout[_gpu] / path_linear_sgd
    (path_sgd_layout.hpp:37-79, path_sgd.cpp:12-31) plus the device knobs."""
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
    n_streams: int = 0
    batch: int = 0
    flags: int = 0
    sampling: int = 0   # 0 auto, 1 stream (reference-exact worker streams), 2 tile
    multi_switch_iteration: int = 0   # hybrid multi-GPU mode: first peer-phase iteration (0 = iter_max // 3)

    def c(self) -> ConfigC:
        return ConfigC(self.iter_max, self.iter_with_max_learning_rate, self.min_term_updates, self.delta, self.eps,
                       self.eta_max, self.theta, self.space, self.space_max, self.space_quantization_step,
                       self.cooling_start, self.seed, self.n_streams, self.batch, self.flags, self.sampling, self.multi_switch_iteration)


@dataclass
class FlatGraph:
    """Host-side flattened graph (path-major SoA), the input of the C-ABI."""
    node_len: np.ndarray
    path_first_step: np.ndarray
    step_node: np.ndarray
    step_rev: Optional[np.ndarray] = None
    step_pos: Optional[np.ndarray] = None
    path_names: list = field(default_factory=list)

    def __post_init__(self):
        self.node_len = np.ascontiguousarray(self.node_len, dtype=np.uint32)
        self.path_first_step = np.ascontiguousarray(self.path_first_step, dtype=np.uint64)
        self.step_node = np.ascontiguousarray(self.step_node, dtype=np.uint32)
        if self.step_rev is not None:
            self.step_rev = np.ascontiguousarray(self.step_rev, dtype=np.uint8)
        if self.step_pos is not None:
            self.step_pos = np.ascontiguousarray(self.step_pos, dtype=np.uint64)

    @property
    def N(self) -> int:
        return int(self.node_len.size)

    @property
    def P(self) -> int:
        return int(self.path_first_step.size - 1)

    @property
    def S(self) -> int:
        return int(self.step_node.size)

    @property
    def max_path_steps(self) -> int:
        return int(np.diff(self.path_first_step.astype(np.int64)).max()) if self.P else 0

    @property
    def max_path_bp(self) -> int:
        """longest path in bp (`odgi sort` uses it as the Zipf space, sort_main.cpp:387)"""
        lens = self.node_len[self.step_node].astype(np.uint64)
        csum = np.concatenate([[0], np.cumsum(lens, dtype=np.uint64)])
        f = self.path_first_step.astype(np.int64)
        return int((csum[f[1:]] - csum[f[:-1]]).max()) if self.P else 0

    def view(self) -> GraphView:
        return GraphView(self.N, self.P, self.S, _ptr(self.node_len), _ptr(self.path_first_step), _ptr(self.step_node),
                         _ptr(self.step_rev), _ptr(self.step_pos))


def layout_defaults(g: FlatGraph, **kw) -> Config:
    """`odgi layout` defaults (layout_main.cpp:198-266): 30 iterations of 10*S updates, eta_max = max_steps^2,
    Zipf space = max path steps, space_max 1000, quantization 100, cooling 0.5."""
    ms = g.max_path_steps
    c = Config(iter_max=30, min_term_updates=10 * g.S, eta_max=float(ms) * float(ms), space=ms, space_max=1000,
               space_quantization_step=100)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def sort_defaults(g: FlatGraph, **kw) -> Config:
    """`odgi sort -Y` defaults (sort_main.cpp:313-414): 100(+1) iterations of 1*S updates, Zipf space = longest
    path in bp, space_max 100, quantization step derived so that about 100 zeta entries exist."""
    ms = g.max_path_steps
    space = g.max_path_bp
    space_max = 100
    max_dists = max(space_max + 1, 100)
    # sort_main.cpp:402-411: the derived step only when space > space_max (else the reference falls back to 100)
    q = max(2, int(np.ceil((space - space_max) / (max_dists - space_max)))) if space > space_max and max_dists > space_max else 100
    c = Config(iter_max=100, min_term_updates=g.S, eta_max=float(ms) * float(ms), space=space, space_max=space_max,
               space_quantization_step=q)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def device_count() -> int:
    return int(lib().pgsgd_device_count())


def schedule(cfg: Config) -> np.ndarray:
    etas = np.zeros(cfg.iter_max + 1, dtype=np.float64)
    cc = cfg.c()
    _check(lib().pgsgd_schedule(C.byref(cc), _ptr(etas)))
    return etas


def zetas(cfg: Config) -> np.ndarray:
    cc = cfg.c()
    n = lib().pgsgd_zetas(C.byref(cc), None, 0)
    z = np.zeros(n, dtype=np.float64)
    lib().pgsgd_zetas(C.byref(cc), _ptr(z), n)
    return z


def scan_gfa(path: str):
    """The host side of the device GFA ingest: line boundaries, node lengths from the S lines (ids must be 1..N), and the byte
    range of every P line's step list.  No per-step work.  Returns (node_len u32[N], file bytes u8[], field_begin u64[P],
    field_end u64[P], path names)."""
    data = np.fromfile(path, dtype=np.uint8)
    nl = np.flatnonzero(data == 10)
    starts = np.concatenate(([0], nl + 1))
    ends = np.concatenate((nl, [data.size]))
    keep = starts < ends
    starts, ends = starts[keep], ends[keep]
    tabs = np.flatnonzero(data == 9)
    first = data[starts]
    # S lines: "S <id> <sequence> ..." -> length of field 3 (or LN:i: when the sequence is '*')
    s_idx = np.flatnonzero(first == ord("S"))
    t1 = np.searchsorted(tabs, starts[s_idx])            # index of the first tab of each S line
    id_b, id_e = tabs[t1] + 1, tabs[t1 + 1]
    seq_b = tabs[t1 + 1] + 1
    t3 = np.minimum(t1 + 2, tabs.size - 1)
    seq_e = np.where((t1 + 2 < tabs.size) & (tabs[t3] < ends[s_idx]), tabs[t3], ends[s_idx])
    ids = np.array([int(bytes(data[b:e])) for b, e in zip(id_b, id_e)], dtype=np.int64) if s_idx.size < 200_000 else None
    if ids is None:   # vectorised decimal parse for big graphs
        ids = np.zeros(s_idx.size, dtype=np.int64)
        width = int((id_e - id_b).max())
        for k in range(width):
            pos = id_e - 1 - k
            ok = pos >= id_b
            ids += np.where(ok, (data[np.where(ok, pos, 0)].astype(np.int64) - 48) * 10 ** k, 0)
    n = int(ids.max()) if ids.size else 0
    if ids.size != n or np.unique(ids).size != n or ids.min() != 1:
        raise PgsgdError(-1, "node ids are not exactly 1..N")
    node_len = np.zeros(n, dtype=np.uint32)
    lens = (seq_e - seq_b).astype(np.uint32)
    star = (lens == 1) & (data[seq_b] == ord("*"))
    for j in np.flatnonzero(star):   # sequence omitted: LN:i: tag
        line = bytes(data[starts[s_idx[j]]:ends[s_idx[j]]])
        k = line.find(b"LN:i:")
        lens[j] = int(line[k + 5:].split(b"\t")[0]) if k >= 0 else 0
    node_len[ids - 1] = lens
    # P lines: "P <name> <steps> ..." -> byte range of field 3
    p_idx = np.flatnonzero(first == ord("P"))
    t1 = np.searchsorted(tabs, starts[p_idx])
    fb = tabs[t1 + 1] + 1
    t3 = np.minimum(t1 + 2, tabs.size - 1)
    fe = np.where((t1 + 2 < tabs.size) & (tabs[t3] < ends[p_idx]), tabs[t3], ends[p_idx])
    fe = np.where((fe > fb) & (data[np.maximum(fe, 1) - 1] == 13), fe - 1, fe)   # CRLF files
    names = [bytes(data[tabs[a] + 1:tabs[a + 1]]).decode() for a in t1]
    return node_len, data, fb.astype(np.uint64), fe.astype(np.uint64), names


class Engine:
    """Device-resident graph + coordinates (pgsgd_engine)."""

    def __init__(self, g: FlatGraph, device: int = 0):
        self.g = g
        self._h = C.c_void_p()
        gv = g.view()
        _check(lib().pgsgd_engine_create(C.byref(gv), device, C.byref(self._h)))

    @classmethod
    def from_gfa(cls, path: str, device: int = 0) -> "Engine":
        """Engine straight from a GFA file: the host only finds the lines (scan_gfa); the P-line step lists are parsed on the
        device (pgsgd_engine_create_from_gfa_paths).  self.g then carries only N, P, S and the node lengths."""
        node_len, data, fb, fe, names = scan_gfa(path)
        e = cls.__new__(cls)
        e._h = C.c_void_p()
        _check(lib().pgsgd_engine_create_from_gfa_paths(_ptr(node_len), len(node_len), _ptr(data), _ptr(fb), _ptr(fe), len(fb), device, C.byref(e._h)))
        st = e.graph_stats()
        e.g = types.SimpleNamespace(N=len(node_len), P=len(fb), S=st["step_count"], node_len=node_len, path_names=names,
                                    max_path_steps=st["max_path_steps"], max_path_bp=st["max_path_bp"])
        return e

    def graph_stats(self) -> dict:
        v = [C.c_uint64(0) for _ in range(4)]
        _check(lib().pgsgd_engine_graph_stats(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("step_count", "max_path_steps", "max_path_bp", "max_node_depth"), (int(x.value) for x in v)))

    def close(self):
        if self._h:
            lib().pgsgd_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def device_bytes(self) -> int:
        return int(lib().pgsgd_engine_device_bytes(self._h))

    def set_coords_2d(self, X, Y):
        X = np.ascontiguousarray(X, dtype=np.float64)
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        assert X.size == 2 * self.g.N and Y.size == 2 * self.g.N
        _check(lib().pgsgd_engine_set_coords_2d(self._h, _ptr(X), _ptr(Y)))

    def get_coords_2d(self, out=None):
        """(X, Y) as float64[2N]; `out` = caller-owned (X, Y) buffers to fill (what the odgi shim does with the caller's vectors)."""
        if out is None:
            X = np.empty(2 * self.g.N, dtype=np.float64)
            Y = np.empty(2 * self.g.N, dtype=np.float64)
        else:
            X, Y = out
            assert X.dtype == np.float64 and Y.dtype == np.float64 and X.size == 2 * self.g.N and Y.size == 2 * self.g.N
            assert X.flags.c_contiguous and Y.flags.c_contiguous
        _check(lib().pgsgd_engine_get_coords_2d(self._h, _ptr(X), _ptr(Y)))
        return X, Y

    def set_coords_2d_f32(self, xy):
        xy = np.ascontiguousarray(xy, dtype=np.float32)
        assert xy.size == 4 * self.g.N
        _check(lib().pgsgd_engine_set_coords_2d_f32(self._h, _ptr(xy)))

    def get_coords_2d_f32(self):
        xy = np.empty(4 * self.g.N, dtype=np.float32)
        _check(lib().pgsgd_engine_get_coords_2d_f32(self._h, _ptr(xy)))
        return xy

    def set_coords_1d(self, X=None):
        if X is not None:
            X = np.ascontiguousarray(X, dtype=np.float64)
            assert X.size == self.g.N
        _check(lib().pgsgd_engine_set_coords_1d(self._h, _ptr(X)))

    def get_coords_1d(self):
        X = np.empty(self.g.N, dtype=np.float64)
        _check(lib().pgsgd_engine_get_coords_1d(self._h, _ptr(X)))
        return X

    def set_frozen_1d(self, frozen):
        f = None if frozen is None else np.ascontiguousarray(frozen, dtype=np.uint8)
        _check(lib().pgsgd_engine_set_frozen_1d(self._h, _ptr(f)))

    def run_2d(self, cfg: Config) -> dict:
        st, cc = StatsC(), cfg.c()
        _check(lib().pgsgd_engine_run_2d(self._h, C.byref(cc), C.byref(st)))
        return st.as_dict()

    def run_1d(self, cfg: Config) -> dict:
        st, cc = StatsC(), cfg.c()
        _check(lib().pgsgd_engine_run_1d(self._h, C.byref(cc), C.byref(st)))
        return st.as_dict()

    def run_range(self, cfg: Config, dims: int, iter_begin: int, iter_end: int) -> dict:
        st, cc = StatsC(), cfg.c()
        _check(lib().pgsgd_engine_run_range(self._h, C.byref(cc), dims, iter_begin, iter_end, C.byref(st)))
        return st.as_dict()

    def attach_comm(self, unique_id: bytes, n_ranks: int, rank: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(lib().pgsgd_engine_attach_comm(self._h, buf, n_ranks, rank))

    def set_trace(self, capacity: int):
        _check(lib().pgsgd_engine_set_trace(self._h, capacity))

    def get_trace(self, capacity: int):
        ia = np.zeros(capacity, np.uint64); ib = np.zeros(capacity, np.uint64); fl = np.zeros(capacity, np.uint8)
        n = C.c_uint64(0)
        _check(lib().pgsgd_engine_get_trace(self._h, _ptr(ia), _ptr(ib), _ptr(fl), C.byref(n)))
        k = int(n.value)
        return ia[:k], ib[:k], fl[:k]

    def path_stress(self, dims: int, n_pairs: int = 1_000_000, seed: int = 12345) -> float:
        out = C.c_double(0.0)
        _check(lib().pgsgd_engine_path_stress(self._h, dims, n_pairs, seed, C.byref(out)))
        return float(out.value)

    def local_stress(self, dims: int, n_pairs: int = 1_000_000, seed: int = 12345) -> float:
        """near-pair stress (partner 1..64 ranks away along the path, <= 1000 bp): pgsgd_engine_local_stress"""
        out = C.c_double(0.0)
        _check(lib().pgsgd_engine_local_stress(self._h, dims, n_pairs, seed, C.byref(out)))
        return float(out.value)

    def order_1d(self, component: Optional[np.ndarray] = None) -> np.ndarray:
        """node ranks sorted by ([component key,] position, handle) on the device (path_sgd.cpp:650-658)"""
        order = np.empty(self.g.N, dtype=np.uint64)
        comp = None if component is None else np.ascontiguousarray(component, dtype=np.uint32)
        _check(lib().pgsgd_engine_order_1d_components(self._h, _ptr(comp), _ptr(order)))
        return order

    def set_multi_mode(self, mode: int):
        """MULTI_ALLREDUCE | MULTI_PEER | MULTI_HYBRID | MULTI_AUTO (include/pgsgd.h)"""
        _check(lib().pgsgd_engine_set_multi_mode(self._h, mode))

    def sort_goodness(self, order: Optional[np.ndarray] = None, gap_links: bool = True, orientation: bool = True) -> dict:
        """`odgi stats -l [-g] -s [-d]` of the graph as `order` would sort it, on the device (pgsgd_engine_sort_goodness)"""
        out = GoodnessC()
        o = None if order is None else np.ascontiguousarray(order, dtype=np.uint64)
        _check(lib().pgsgd_engine_sort_goodness(self._h, _ptr(o), (1 if gap_links else 0) | (2 if orientation else 0), C.byref(out)))
        return {f: getattr(out, f) for f, _ in GoodnessC._fields_}

    def encode_lay(self, component: Optional[np.ndarray] = None) -> bytes:
        """the resident 2D layout as odgi's .lay file, encoded on the device (pgsgd_engine_encode_lay); component: [N] weak
        component id per node (the per-component stacking of `odgi layout` is applied first) or None"""
        comp = None if component is None else np.ascontiguousarray(component, dtype=np.uint32)
        k = 0 if comp is None else int(comp.max()) + 1
        n = C.c_uint64(0)
        _check(lib().pgsgd_engine_encode_lay(self._h, _ptr(comp), k, None, 0, C.byref(n)))
        buf = np.empty(n.value, dtype=np.uint8)
        _check(lib().pgsgd_engine_encode_lay(self._h, _ptr(comp), k, _ptr(buf), n.value, C.byref(n)))
        return buf.tobytes()

    def resolved_multi_mode(self) -> int:
        """MULTI_* in effect (what MULTI_AUTO resolved to once the coordinates were uploaded)"""
        return int(lib().pgsgd_engine_resolved_multi_mode(self._h))

    def set_shard(self, global_step_count: int):
        """this engine holds only some of the job's paths (graphio.partition_paths); 0 switches the mode off"""
        _check(lib().pgsgd_engine_set_shard(self._h, int(global_step_count)))

    def sample_terms(self, cfg: Config, dims: int, cooling: bool, n_terms: int, stream: int = 0, theta_zipf=None):
        out = {"step_index": np.zeros(n_terms, np.uint64), "path": np.zeros(n_terms, np.uint32),
               "rank_a": np.zeros(n_terms, np.uint64), "rank_b": np.zeros(n_terms, np.uint64),
               "node_a": np.zeros(n_terms, np.uint32), "node_b": np.zeros(n_terms, np.uint32),
               "pos_a": np.zeros(n_terms, np.uint64), "pos_b": np.zeros(n_terms, np.uint64),
               "end_a": np.zeros(n_terms, np.uint8), "end_b": np.zeros(n_terms, np.uint8),
               "valid": np.zeros(n_terms, np.uint8)}
        cc = cfg.c()
        th = cfg.theta if theta_zipf is None else theta_zipf
        _check(lib().pgsgd_engine_sample_terms(self._h, C.byref(cc), dims, int(cooling), th, stream, n_terms,
                                               *[_ptr(out[k]) for k in ("step_index", "path", "rank_a", "rank_b", "node_a",
                                                                        "node_b", "pos_a", "pos_b", "end_a", "end_b", "valid")]))
        return out


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    _check(lib().pgsgd_comm_unique_id(buf))
    return bytes(buf)


def layout_2d(g: FlatGraph, cfg: Config, X, Y):
    """One-shot pgsgd_layout_2d: host buffers in/out (the call odgi's shim makes)."""
    X = np.ascontiguousarray(X, dtype=np.float64).copy()
    Y = np.ascontiguousarray(Y, dtype=np.float64).copy()
    st, cc, gv = StatsC(), cfg.c(), g.view()
    _check(lib().pgsgd_layout_2d(C.byref(gv), C.byref(cc), _ptr(X), _ptr(Y), C.byref(st)))
    return X, Y, st.as_dict()


def layout_2d_multi(g: FlatGraph, cfg: Config, X, Y, n_gpus: int, multi_mode: int = 2):
    """pgsgd_layout_2d_multi: one process, one host thread per GPU inside the call"""
    X = np.ascontiguousarray(X, dtype=np.float64).copy()
    Y = np.ascontiguousarray(Y, dtype=np.float64).copy()
    st, cc, gv = StatsC(), cfg.c(), g.view()
    _check(lib().pgsgd_layout_2d_multi(C.byref(gv), C.byref(cc), n_gpus, multi_mode, _ptr(X), _ptr(Y), C.byref(st)))
    return X, Y, st.as_dict()


def sort_1d(g: FlatGraph, cfg: Config, X=None, frozen=None):
    """One-shot pgsgd_sort_1d."""
    init = X is not None
    X = np.ascontiguousarray(X, dtype=np.float64).copy() if init else np.zeros(g.N, dtype=np.float64)
    f = None if frozen is None else np.ascontiguousarray(frozen, dtype=np.uint8)
    st, cc, gv = StatsC(), cfg.c(), g.view()
    _check(lib().pgsgd_sort_1d(C.byref(gv), C.byref(cc), _ptr(f), int(init), _ptr(X), C.byref(st)))
    return X, st.as_dict()
