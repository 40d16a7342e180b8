"""The oracle (oracle/pgsgd_oracle.c) against golden vectors produced by the reference itself.

The fixtures were written by scripts/pin_oracle.py from runs of the UNMODIFIED reference
(oracle/_ref/ref_driver_trace): per-term traces via the reference's own -Deval_path_sgd hook and the
final fp64 coordinates of single-thread runs.  Everything here is bit-exact.
"""
import json
import os

import numpy as np
import pytest

from odgi_b200.arrays import read_arrays
from oracle import oracle as orc

UNKNOWN_PATH = 0xFFFFFFFF


def _load(golden_dir, name):
    return read_arrays(os.path.join(golden_dir, name))


def test_pathindex_known_answer():
    """The reference's own KAT for the integer half (src/unittest/pathindex.cpp:20-131): a 4-node,
    3-path graph whose step positions must be 0,0,4,6 for path "5"'s steps 1+,3+,4+ ... here
    restated on the note5-like graph of that test: node lengths 4,1,2,7; path 1+,3+,4+."""
    node_len = np.array([4, 1, 2, 7], dtype=np.uint32)
    first = np.array([0, 3, 6], dtype=np.uint64)
    step_node = np.array([0, 2, 3, 0, 2, 3], dtype=np.uint32)
    pos = orc.positions_from_lengths(node_len, first, step_node)
    assert pos.tolist() == [0, 4, 6, 0, 4, 6]  # pathindex.cpp:126-131 checks 0,4,6 per path


def test_flatten_matches_xp(golden_graphs):
    for name, a in golden_graphs.items():
        pos = orc.positions_from_lengths(a["node_len"], a["path_first_step"], a["step_node"])
        assert np.array_equal(pos, a["step_pos"]), name  # step_pos was asserted == XP get_position_of_step at dump time
        # XP's node-major table is a permutation of all (path, rank) pairs
        g = orc.Graph.from_arrays(a, use_xp_perm=True)
        assert np.array_equal(np.sort(g.step_perm), np.arange(g.S, dtype=np.uint64)), name
        # ... grouped by node in ascending node order (xp.cpp:127-141)
        assert np.all(np.diff(g.step_node[g.step_perm].astype(np.int64)) >= 0), name


def test_schedule_bit_exact(golden_dir):
    with open(os.path.join(golden_dir, "schedule.json")) as f:
        cases = json.load(f)
    for c in cases:
        ref = np.array([float.fromhex(x) for x in c["etas_hex"]])
        mine = orc.schedule(orc.Config(iter_max=c["iter_max"], iter_with_max_learning_rate=c["iter_lr"], eta_max=c["eta_max"], eps=c["eps"]))
        assert np.array_equal(ref, mine)


@pytest.mark.parametrize("name,tag", [("DRB1-3123", "nocool"), ("DRB1-3123", "cool"), ("chr6.C4", "nocool"), ("chr6.C4", "cool"), ("LPA", "cool"),
                                      # odd shapes: single-step path + back-to-back repeat, two short paths, a reverse-strand step
                                      ("overlap", "cool"), ("k", "cool"), ("note5", "cool")])
def test_2d_trace_and_coords_bit_exact(golden_dir, golden_graphs, name, tag):
    pin = _load(golden_dir, f"{name}.pin2d_{tag}.arr.gz")
    g = orc.Graph.from_arrays(golden_graphs[name], use_xp_perm=True)
    eta = float(pin["eta"][0])
    ms = g.max_path_steps
    cfg = orc.Config(iter_max=2, min_term_updates=int(pin["updates"][0]), eps=eta, eta_max=eta, theta=0.99, space=ms,
                     space_max=1000, space_quantization_step=100, cooling_start=float(pin["cooling_start"][0]))
    X, Y = orc.layout_init(g, seed=int(pin["init_seed"][0]))
    n = len(pin["trace_pos_a"])
    terms = orc.replay_single(g, cfg, 2, n, int(pin["switch_at"][0]), eta, eta, False, True, 0.99, X, Y)
    assert np.array_equal(terms["pos_a"], pin["trace_pos_a"])
    assert np.array_equal(terms["pos_b"], pin["trace_pos_b"])
    tp = pin["trace_path"]
    assert np.all((terms["path"] == tp) | (tp == UNKNOWN_PATH))
    assert np.array_equal(X, pin["X"]) and np.array_equal(Y, pin["Y"])


@pytest.mark.parametrize("name", ["DRB1-3123", "LPA", "chr6.C4", "overlap", "k", "note5"])
def test_1d_trace_and_coords_bit_exact(golden_dir, golden_graphs, name):
    pin = _load(golden_dir, f"{name}.pin1d.arr.gz")
    g = orc.Graph.from_arrays(golden_graphs[name], use_xp_perm=True)
    eta = float(pin["eta"][0])
    cfg = orc.default_sort_config(g, iter_max=2, min_term_updates=int(pin["updates"][0]), eps=eta, eta_max=eta, cooling_start=0.0)
    X = orc.sort_init(g)
    n = len(pin["trace_pos_a"])
    terms = orc.replay_single(g, cfg, 1, n, int(pin["switch_at"][0]), eta, eta, False, True, 0.001, X, None)
    assert np.array_equal(terms["pos_a"], pin["trace_pos_a"])
    assert np.array_equal(terms["pos_b"], pin["trace_pos_b"])
    assert np.array_equal(X, pin["X"])


@pytest.mark.parametrize("name,mod", [("DRB1-3123", 3), ("LPA", 2)])
def test_1d_frozen_nodes_bit_exact(golden_dir, golden_graphs, name, mod):
    """`odgi sort -H` semantics (target nodes stay put, path_sgd.cpp:290-302,387-392): reference run with every third
    (second) node frozen, replayed by the oracle — trace and final coordinates bit-exact; frozen nodes never moved."""
    pin = _load(golden_dir, f"{name}.pin1d_frozen{mod}.arr.gz")
    g = orc.Graph.from_arrays(golden_graphs[name], use_xp_perm=True)
    mod = int(pin["freeze_mod"][0])
    frozen = (np.arange(g.N) % mod == 0).astype(np.uint8)
    eta = float(pin["eta"][0])
    cfg = orc.default_sort_config(g, iter_max=2, min_term_updates=int(pin["updates"][0]), eps=eta, eta_max=eta, cooling_start=0.0)
    X = orc.sort_init(g)
    n = len(pin["trace_pos_a"])
    terms = orc.replay_single(g, cfg, 1, n, int(pin["switch_at"][0]), eta, eta, False, True, 0.001, X, None, frozen=frozen)
    assert np.array_equal(terms["pos_a"], pin["trace_pos_a"]) and np.array_equal(terms["pos_b"], pin["trace_pos_b"])
    assert np.array_equal(X, pin["X"])
    assert np.array_equal(X[frozen == 1], orc.sort_init(g)[frozen == 1])


def test_zipf_range_and_edge_cases():
    L = orc.lib()
    import ctypes as C
    rng = orc._Rng()
    L.orc_rng_seed(C.byref(rng), 1)
    for theta, ztheta in ((0.99, 0.99), (0.001, 0.99)):
        for n in (1, 2, 3, 4, 10, 1000, 3100):
            zn = L.orc_zeta(n, ztheta)
            draws = [L.orc_dirty_zipf(C.byref(rng), n, theta, zn) for _ in range(2000)]
            assert min(draws) >= 1 and max(draws) <= n, (theta, n, min(draws), max(draws))


def test_runs_are_deterministic_and_count_updates(golden_graphs):
    g = orc.Graph.from_arrays(golden_graphs["DRB1-3123"])
    cfg = orc.default_layout_config(g, iter_max=3, min_term_updates=5000)
    X0, Y0 = orc.layout_init(g, 7)
    n1, X1, Y1 = orc.layout_2d(g, cfg, X0.copy(), Y0.copy(), n_streams=4)
    n2, X2, Y2 = orc.layout_2d(g, cfg, X0.copy(), Y0.copy(), n_streams=4)
    assert n1 == n2 == 3 * 5000
    assert np.array_equal(X1, X2) and np.array_equal(Y1, Y2)
    cfg1 = orc.default_sort_config(g, iter_max=3, min_term_updates=5000)
    n3, _ = orc.sort_1d(g, cfg1, orc.sort_init(g), n_streams=3)
    assert n3 == 4 * 5000  # 1D runs iter_max + 1 iterations (path_sgd.cpp:181)


def test_staleness_model_reduces_to_the_sequential_run(golden_graphs):
    """orc_run_inflight (planning tool for the launch caps): one term in flight is the sequential single-stream run up to
    fp32 rounding of the intermediate products; a few hundred in flight stay in the reference's stress band."""
    g = orc.Graph.from_arrays(golden_graphs["DRB1-3123"])
    cfg = orc.default_layout_config(g, iter_max=6, min_term_updates=20000)
    X0, Y0 = orc.layout_init(g, 42)
    _, seq = orc.layout_2d_f32(g, cfg, orc.XY_to_xy(X0, Y0), n_streams=1)
    xy = orc.XY_to_xy(X0, Y0)
    assert orc.run_inflight(g, cfg, 1, 2, False, xy=xy) == 6 * 20000
    s_seq = orc.path_stress_2d(g, *orc.xy_to_XY(seq), 200000, 1)
    s_one = orc.path_stress_2d(g, *orc.xy_to_XY(xy), 200000, 1)
    assert abs(s_one - s_seq) <= 0.02 * s_seq, (s_one, s_seq)
    x = orc.sort_init(g)
    c1 = orc.default_sort_config(g, iter_max=10)
    assert orc.run_inflight(g, c1, 256, 1, True, X=x) == 11 * c1.min_term_updates and np.all(np.isfinite(x))
