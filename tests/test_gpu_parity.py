"""Parity of the CUDA path (through the C-ABI) against the oracle, on a real B200.

Levels, strongest first:
  1. sampler: the device draws the SAME terms as a reference worker thread with the same seed — bit-exact on every
     integer field (step index, path, ranks, nodes, end choice, end-adjusted bp positions);
  2. arithmetic: with one worker stream and batch 1 (strict order) the device coordinates equal the oracle's
     fp32 device model (2D) / the reference's fp64 arithmetic (1D) bit-for-bit;
  3. Hogwild runs: sampled path stress of full default runs within tolerance of the reference CPU implementation.
"""
import json
import os

import numpy as np
import pytest

import odgi_b200
from odgi_b200 import capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

INT_FIELDS = ("step_index", "path", "rank_a", "rank_b", "node_a", "node_b", "pos_a", "pos_b", "end_a", "end_b")


@pytest.fixture(scope="module")
def graphs(golden_graphs):
    return {k: (odgi_b200.graph_from_arrays(a), orc.Graph.from_arrays(a)) for k, a in golden_graphs.items()}


def _cfgs(gd, go, dims, **kw):
    if dims == 2:
        return capi.layout_defaults(gd, **kw), orc.default_layout_config(go, **kw)
    return capi.sort_defaults(gd, **kw), orc.default_sort_config(go, **kw)


@pytest.mark.parametrize("name,dims,cooling,theta_zipf", [
    ("DRB1-3123", 2, False, None), ("DRB1-3123", 2, True, None), ("chr6.C4", 2, False, None), ("chr6.C4", 2, True, None),
    ("LPA", 1, False, None), ("LPA", 1, True, 0.001), ("DRB1-3123", 1, True, 0.001), ("note5", 2, False, None), ("t", 2, True, None)])
def test_sampler_bit_exact(graphs, name, dims, cooling, theta_zipf):
    gd, go = graphs[name]
    cd, co = _cfgs(gd, go, dims)
    n = 20000
    with odgi_b200.Engine(gd) as e:
        for stream in (0, 1, 77777):
            dev = e.sample_terms(cd, dims, cooling, n, stream=stream, theta_zipf=theta_zipf)
            ref, valid = orc.sample_terms(go, co, dims, cooling, n, stream=stream, theta_zipf=theta_zipf)
            assert np.array_equal(dev["valid"], valid)
            ok = valid.astype(bool)
            for f in INT_FIELDS:
                if dims == 1 and f in ("end_a", "end_b"):
                    continue
                assert np.array_equal(dev[f][ok], ref[f][ok].astype(dev[f].dtype)), (name, stream, f)
            # the integer path distance itself
            assert np.array_equal(np.abs(dev["pos_a"][ok].astype(np.int64) - dev["pos_b"][ok].astype(np.int64)),
                                  np.abs(ref["pos_a"][ok].astype(np.int64) - ref["pos_b"][ok].astype(np.int64)))


def test_sampler_one_step_paths():
    """Paths of a single step are drawn but not counted (path_sgd_layout.cpp:190-192); sort.cpp:129-255 is the
    reference's own test of that corner."""
    node_len = np.array([3, 1, 2, 5, 4], dtype=np.uint32)
    first = np.array([0, 1, 4, 5, 5, 7], dtype=np.uint64)  # a 1-step, a 3-step, a 1-step, an EMPTY and a 2-step path
    step_node = np.array([0, 1, 2, 3, 4, 0, 4], dtype=np.uint32)
    step_rev = np.array([0, 1, 0, 0, 1, 0, 0], dtype=np.uint8)
    gd = odgi_b200.FlatGraph(node_len, first, step_node, step_rev)
    go = orc.Graph(node_len, first, step_node, step_rev)
    cd, co = _cfgs(gd, go, 2)
    with odgi_b200.Engine(gd) as e:
        dev = e.sample_terms(cd, 2, False, 5000)
    ref, valid = orc.sample_terms(go, co, 2, False, 5000)
    assert np.array_equal(dev["valid"], valid) and 0 < valid.sum() < 5000
    ok = valid.astype(bool)
    for f in INT_FIELDS:
        assert np.array_equal(dev[f][ok], ref[f][ok].astype(dev[f].dtype)), f


@pytest.mark.parametrize("flags", [0, capi.PGSGD_FLAG_EXCH_WRITE, capi.PGSGD_FLAG_PLAIN_STORE])
@pytest.mark.parametrize("name", ["DRB1-3123", "chr6.C4"])
def test_2d_single_stream_bit_exact(graphs, name, flags):
    gd, go = graphs[name]
    kw = dict(iter_max=4, min_term_updates=6000, eta_max=2000.0)
    cd, co = _cfgs(gd, go, 2, **kw)
    cd.n_streams, cd.batch, cd.flags = 1, 1, flags
    X0, Y0 = orc.layout_init(go, seed=3)
    xy0 = orc.XY_to_xy(X0, Y0)
    n_ref, xy_ref = orc.layout_2d_f32(go, co, xy0.copy(), n_streams=1)
    with odgi_b200.Engine(gd) as e:
        e.set_coords_2d_f32(xy0)
        st = e.run_2d(cd)
        xy_dev = e.get_coords_2d_f32()
    assert st["term_updates"] == n_ref == 4 * 6000
    assert np.array_equal(xy_dev, xy_ref)


@pytest.mark.parametrize("flags", [0, capi.PGSGD_FLAG_EXCH_WRITE])
@pytest.mark.parametrize("name", ["LPA", "DRB1-3123"])
def test_1d_single_stream_bit_exact(graphs, name, flags):
    gd, go = graphs[name]
    kw = dict(iter_max=3, min_term_updates=5000, eta_max=2000.0, cooling_start=0.3)
    cd, co = _cfgs(gd, go, 1, **kw)
    cd.n_streams, cd.batch, cd.flags = 1, 1, flags
    n_ref, x_ref = orc.sort_1d(go, co, orc.sort_init(go), n_streams=1)
    with odgi_b200.Engine(gd) as e:
        e.set_coords_1d(None)
        st = e.run_1d(cd)
        x_dev = e.get_coords_1d()
    assert st["term_updates"] == n_ref == 4 * 5000  # iter_max + 1 iterations
    assert np.array_equal(x_dev, x_ref)


def test_1d_frozen_nodes_stay_put(graphs):
    gd, go = graphs["DRB1-3123"]
    cd, co = _cfgs(gd, go, 1, iter_max=5)
    frozen = (np.arange(gd.N) % 3 == 0).astype(np.uint8)
    x0 = orc.sort_init(go)
    x, st = odgi_b200.sort_1d(gd, cd, x0, frozen=frozen)
    assert np.array_equal(x[frozen == 1], x0[frozen == 1])
    assert np.any(x[frozen == 0] != x0[frozen == 0])
    assert st["term_updates"] == 6 * cd.min_term_updates
    # single stream: identical to the oracle with the same frozen set
    cd.n_streams, cd.batch = 1, 1
    cd.min_term_updates = co.min_term_updates = 4000
    n_ref, x_ref = orc.sort_1d(go, co, x0.copy(), n_streams=1, frozen=frozen)
    x1, _ = odgi_b200.sort_1d(gd, cd, x0, frozen=frozen)
    assert np.array_equal(x1, x_ref)


def _stress_band(golden_dir, key):
    with open(os.path.join(golden_dir, "stress_reference.json")) as f:
        return json.load(f)[key]


def _assert_in_band(values, band):
    """|stress_gpu - mean(stress_cpu)| <= max(1 % of the reference mean, 2 sigma_cpu), applied to the mean over
    the seeds.  A SINGLE run drawn from the reference's own distribution leaves a 2-sigma band one time in ~16 (the reference
    mean itself comes from six runs), so the per-run gate is max(1 %, 3 sigma): the same bound, one sigma wider.
    The band = six runs of the reference (hard-coded worker seeds: thread timing only) + twelve runs of the oracle under other
    seeds (scripts/make_stress_golden.py --oracle-seeds); with the reference runs alone a seed whose OWN expectation sits
    0.9 % off their mean (DRB1-3123 2D, seed 42: oracle 0.06886, GPU 0.0688, reference runs 0.06825) is one noisy run away
    from the 1 % gate."""
    mean, sd = band["mean"], band["sd"]
    assert abs(np.mean(values) - mean) <= max(0.01 * mean, 2 * sd), (values, mean, sd)
    assert all(abs(v - mean) <= max(0.01 * mean, 3 * sd) for v in values), (values, mean, sd)


@pytest.mark.parametrize("name", ["DRB1-3123", "chr6.C4"])
def test_2d_default_run_stress_within_reference_band(graphs, golden_dir, name):
    """Full `odgi layout` default runs (30 x 10*S updates, Hogwild, automatic launch shape) vs the reference CPU
    implementation: sampled path stress within max(1 %, 2 sigma) of the mean over the reference's own runs
    (tests/golden/stress_reference.json, written by scripts/make_stress_golden.py from oracle/_ref runs, same init)."""
    gd, go = graphs[name]
    band = _stress_band(golden_dir, f"{name}.layout2d")
    X0, Y0 = orc.layout_init(go, seed=42)
    vals = []
    for seed in (9399220, 1234567, 42):
        cd = capi.layout_defaults(gd, seed=seed)
        X, Y, st = odgi_b200.layout_2d(gd, cd, X0, Y0)
        assert st["term_updates"] == 30 * 10 * gd.S
        assert np.all(np.isfinite(X)) and np.all(np.isfinite(Y))
        vals.append(orc.path_stress_2d(go, X, Y, n_pairs=band["n_pairs"], seed=band["seed"]))
    _assert_in_band(vals, band)


@pytest.mark.parametrize("name", ["LPA", "DRB1-3123"])
def test_1d_default_run_stress_within_reference_band(graphs, golden_dir, name):
    gd, go = graphs[name]
    band = _stress_band(golden_dir, f"{name}.sort1d")
    vals = []
    for seed in (9399220, 1234567, 42):
        cd = capi.sort_defaults(gd, seed=seed)
        x, st = odgi_b200.sort_1d(gd, cd)
        assert st["term_updates"] == 101 * gd.S
        vals.append(orc.path_stress_1d(go, x, n_pairs=band["n_pairs"], seed=band["seed"]))
        # the node order derived from X is a permutation
        order = orc.order_from_x(x)
        assert np.array_equal(np.sort(order), np.arange(gd.N, dtype=np.uint64))
    _assert_in_band(vals, band)


def test_engine_equals_one_shot(graphs):
    gd, go = graphs["DRB1-3123"]
    cd = capi.layout_defaults(gd, iter_max=3, n_streams=1, batch=1, min_term_updates=3000)
    X0, Y0 = orc.layout_init(go, seed=9)
    X1, Y1, _ = odgi_b200.layout_2d(gd, cd, X0, Y0)
    with odgi_b200.Engine(gd) as e:
        e.set_coords_2d(X0, Y0)
        e.run_2d(cd)
        X2, Y2 = e.get_coords_2d()
    assert np.array_equal(X1, X2) and np.array_equal(Y1, Y2)


def test_delta_early_stop(graphs):
    gd, go = graphs["DRB1-3123"]
    cd = capi.layout_defaults(gd, delta=1e30)  # every iteration's max |Delta| is below this: stop after the first
    X0, Y0 = orc.layout_init(go, seed=1)
    _, _, st = odgi_b200.layout_2d(gd, cd, X0, Y0)
    assert st["iterations_run"] == 1 and st["last_delta_max"] > 0


@pytest.mark.parametrize("flags", [capi.PGSGD_FLAG_EXCH_WRITE, capi.PGSGD_FLAG_PLAIN_STORE])
def test_write_variants_run(graphs, flags):
    gd, go = graphs["DRB1-3123"]
    cd = capi.layout_defaults(gd, flags=flags)
    X0, Y0 = orc.layout_init(go, seed=42)
    X, Y, st = odgi_b200.layout_2d(gd, cd, X0, Y0)
    s0 = orc.path_stress_2d(go, X0, Y0, 200000, 5)
    s1 = orc.path_stress_2d(go, X, Y, 200000, 5)
    assert np.all(np.isfinite(X)) and s1 < s0


# ---- tile sampling (PGSGD_SAMPLING_TILE): same partner law, stratified first step -------------------------------

@pytest.mark.parametrize("name", ["DRB1-3123", "chr6.C4"])
def test_2d_tile_sampling_stress_within_reference_band(graphs, golden_dir, name):
    gd, go = graphs[name]
    band = _stress_band(golden_dir, f"{name}.layout2d")
    X0, Y0 = orc.layout_init(go, seed=42)
    vals = []
    for seed in (9399220, 1234567, 42):
        cd = capi.layout_defaults(gd, seed=seed, sampling=capi.SAMPLING_TILE)
        X, Y, st = odgi_b200.layout_2d(gd, cd, X0, Y0)
        # every step is a first step exactly 10 times per iteration
        assert st["term_updates"] == 30 * 10 * gd.S
        assert np.all(np.isfinite(X)) and np.all(np.isfinite(Y))
        vals.append(orc.path_stress_2d(go, X, Y, n_pairs=band["n_pairs"], seed=band["seed"]))
    _assert_in_band(vals, band)


@pytest.mark.parametrize("name", ["LPA", "DRB1-3123"])
def test_1d_tile_sampling_stress_within_reference_band(graphs, golden_dir, name):
    gd, go = graphs[name]
    band = _stress_band(golden_dir, f"{name}.sort1d")
    vals = []
    for seed in (9399220, 1234567, 42):
        cd = capi.sort_defaults(gd, seed=seed, sampling=capi.SAMPLING_TILE)
        x, st = odgi_b200.sort_1d(gd, cd)
        # 1D skips terms of zero path distance without counting them (path_sgd.cpp:320-323)
        assert 0.97 * 101 * gd.S <= st["term_updates"] <= 101 * gd.S
        vals.append(orc.path_stress_1d(go, x, n_pairs=band["n_pairs"], seed=band["seed"]))
    _assert_in_band(vals, band)


def test_tile_tma_staging_variant(graphs, golden_dir):
    """PGSGD_FLAG_TMA_STAGING: tiles staged by double-buffered TMA bulk copies + mbarrier; same results statistically"""
    gd, go = graphs["chr6.C4"]
    band = _stress_band(golden_dir, "chr6.C4.layout2d")
    X0, Y0 = orc.layout_init(go, seed=42)
    cd = capi.layout_defaults(gd, sampling=capi.SAMPLING_TILE, flags=capi.PGSGD_FLAG_TMA_STAGING)
    X, Y, st = odgi_b200.layout_2d(gd, cd, X0, Y0)
    assert st["term_updates"] == 30 * 10 * gd.S
    s = orc.path_stress_2d(go, X, Y, n_pairs=band["n_pairs"], seed=band["seed"])
    assert abs(s - band["mean"]) <= 0.025 * band["mean"], (s, band["mean"])


def test_tile_sampling_partial_pass_counts(graphs):
    """U not a multiple of S: full passes + a truncated one; the counted updates are exact when the truncated pass
    does not reach the short last tile"""
    gd, go = graphs["chr6.C4"]
    U = 3 * gd.S + 5000
    cd = capi.layout_defaults(gd, iter_max=2, min_term_updates=U, sampling=capi.SAMPLING_TILE)
    X0, Y0 = orc.layout_init(go, seed=1)
    _, _, st = odgi_b200.layout_2d(gd, cd, X0, Y0)
    assert abs(int(st["term_updates"]) - 2 * U) <= 2 * 2048


def _jump_hist(rank_a, rank_b, edges):
    d = rank_b.astype(np.int64) - rank_a.astype(np.int64)
    return np.histogram(np.abs(d), bins=edges)[0].astype(np.float64), float(np.mean(d < 0))


@pytest.mark.parametrize("name,dims,cooling", [("chr6.C4", 2, False), ("chr6.C4", 2, True), ("LPA", 1, True), ("DRB1-3123", 2, True)])
def test_tile_sampler_law_matches_reference_law(graphs, name, dims, cooling):
    """The tile kernel's economical partner draw (one RNG word per term, fp32 dirty Zipf) against the oracle's exact
    reference sampler: every step is a first step exactly once per pass, and the jump-length histogram, direction,
    Zipf-vs-uniform and end-flip frequencies agree (two-sample chi-square on log-spaced jump bins)."""
    gd, go = graphs[name]
    iter_max = 4
    kw = dict(iter_max=iter_max, min_term_updates=gd.S, sampling=capi.SAMPLING_TILE, cooling_start=0.5)
    cd = capi.layout_defaults(gd, **kw) if dims == 2 else capi.sort_defaults(gd, **kw)
    it = 3 if cooling else 0   # 2D cools from iteration 2, 1D from iteration 3
    with odgi_b200.Engine(gd) as e:
        if dims == 2:
            e.set_coords_2d(*orc.layout_init(go, 1))
        else:
            e.set_coords_1d(None)
        if it:
            e.run_range(cd, dims, 0, it)
        e.set_trace(gd.S + 16)
        e.run_range(cd, dims, it, it + 1)
        ia, ib, fl = e.get_trace(gd.S + 16)
    first = gd.path_first_step.astype(np.int64)
    counts = np.diff(first)
    multi = np.repeat(counts > 1, counts)
    # every step of a multi-step path is a first step exactly once
    assert ia.size == int(multi.sum())
    assert np.array_equal(np.sort(ia), np.nonzero(multi)[0].astype(np.uint64))
    pa = np.searchsorted(first, ia.astype(np.int64), side="right") - 1
    pb = np.searchsorted(first, ib.astype(np.int64), side="right") - 1
    assert np.array_equal(pa, pb)  # partners never leave the path
    ra, rb = ia.astype(np.int64) - first[pa], ib.astype(np.int64) - first[pb]
    # reference law from the oracle (same graph, same config), ~the same number of terms
    co = (orc.default_layout_config(go, iter_max=iter_max, min_term_updates=gd.S) if dims == 2
          else orc.default_sort_config(go, iter_max=iter_max, min_term_updates=gd.S))
    theta_z = 0.001 if (dims == 1 and cooling) else None
    n_ref = min(int(ia.size), 200_000)
    ref, valid = orc.sample_terms(go, co, dims, cooling, n_ref, stream=5, theta_zipf=theta_z)
    ok = valid.astype(bool)
    edges = np.unique(np.concatenate([[0, 1, 2, 3, 4, 6, 8, 12, 16], np.round(np.geomspace(24, max(int(counts.max()), 32), 24)).astype(np.int64)]))
    edges = np.append(edges, edges[-1] + 1)
    h_dev, back_dev = _jump_hist(ra, rb, edges)
    h_ref, back_ref = _jump_hist(ref["rank_a"][ok], ref["rank_b"][ok], edges)
    keep = (h_dev + h_ref) >= 20
    a, b = h_dev[keep], h_ref[keep]
    k1, k2 = np.sqrt(b.sum() / a.sum()), np.sqrt(a.sum() / b.sum())
    chi2 = float((((k1 * a - k2 * b) ** 2) / (a + b)).sum())
    dof = int(keep.sum()) - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof, h_dev, h_ref)
    assert abs(back_dev - back_ref) < 0.01
    if dims == 2:
        assert abs(np.mean(fl & 1) - 0.5) < 0.01 and abs(np.mean(fl >> 1) - 0.5) < 0.01


def test_device_stress_equals_oracle_stress(graphs):
    """pgsgd_engine_path_stress evaluates the oracle's definition with the same 4096 generators: identical pair sets,
    IEEE fp64 arithmetic — equal up to the order of the final 4096-term sum (here: bit-identical)."""
    gd, go = graphs["chr6.C4"]
    X0, Y0 = orc.layout_init(go, 5)
    with odgi_b200.Engine(gd) as e:
        e.set_coords_2d(X0, Y0)
        s_dev0 = e.path_stress(2, 300_000, 99)
        e.run_2d(capi.layout_defaults(gd, iter_max=6))
        X, Y = e.get_coords_2d()
        s_dev = e.path_stress(2, 300_000, 99)
    assert s_dev0 == orc.path_stress_2d(go, X0.astype(np.float32), Y0.astype(np.float32), 300_000, 99)
    assert s_dev == orc.path_stress_2d(go, X, Y, 300_000, 99)
    gd, go = graphs["LPA"]
    with odgi_b200.Engine(gd) as e:
        e.set_coords_1d(None)
        e.run_1d(capi.sort_defaults(gd, iter_max=5))
        x = e.get_coords_1d()
        s1 = e.path_stress(1, 200_000, 7)
    assert s1 == orc.path_stress_1d(go, x, 200_000, 7)


def test_device_derived_positions_and_id_validation(golden_graphs):
    """Without caller-supplied step_pos the engine derives the positions on the device (scan of node lengths): the
    sampler must still produce XP's positions bit for bit; a step pointing past the node table is rejected like the
    reference rejects non-compacted ids (layout.cu:320-323)."""
    a = golden_graphs["chr6.C4"]
    gd = odgi_b200.FlatGraph(a["node_len"], a["path_first_step"], a["step_node"], a["step_rev"], None)
    go = orc.Graph.from_arrays(a)
    cd, co = capi.layout_defaults(gd), orc.default_layout_config(go)
    with odgi_b200.Engine(gd) as e:
        dev = e.sample_terms(cd, 2, True, 20000, stream=3)
    ref, valid = orc.sample_terms(go, co, 2, True, 20000, stream=3)
    ok = valid.astype(bool)
    for f in ("pos_a", "pos_b", "node_a", "node_b", "rank_a", "rank_b"):
        assert np.array_equal(dev[f][ok], ref[f][ok].astype(dev[f].dtype)), f
    bad = a["step_node"].copy()
    bad[1234] = a["node_len"].size + 5
    with pytest.raises(odgi_b200.PgsgdError) as ei:
        odgi_b200.Engine(odgi_b200.FlatGraph(a["node_len"], a["path_first_step"], bad, a["step_rev"], None))
    assert ei.value.code == -6


def test_many_paths_use_the_global_path_table():
    """More paths than the shared-memory table holds (> 25 k): both kernels fall back to the global path_first table."""
    rng = np.random.Generator(np.random.PCG64(3))
    P, N = 30_000, 5_000
    counts = rng.integers(1, 6, size=P)
    first = np.zeros(P + 1, dtype=np.uint64)
    np.cumsum(counts, out=first[1:])
    S = int(first[-1])
    node_len = rng.integers(1, 30, size=N).astype(np.uint32)
    step_node = rng.integers(0, N, size=S).astype(np.uint32)
    step_rev = rng.integers(0, 2, size=S).astype(np.uint8)
    gd = odgi_b200.FlatGraph(node_len, first, step_node, step_rev)
    go = orc.Graph(node_len, first, step_node, step_rev)
    cd, co = capi.layout_defaults(gd), orc.default_layout_config(go)
    with odgi_b200.Engine(gd) as e:
        dev = e.sample_terms(cd, 2, False, 8000, stream=1)
        ref, valid = orc.sample_terms(go, co, 2, False, 8000, stream=1)
        assert np.array_equal(dev["valid"], valid)
        ok = valid.astype(bool)
        for f in INT_FIELDS:
            assert np.array_equal(dev[f][ok], ref[f][ok].astype(dev[f].dtype)), f
        X0, Y0 = orc.layout_init(go, 2)
        s0 = orc.path_stress_2d(go, X0, Y0, 200000, 1)
        for sampling in (capi.SAMPLING_STREAM, capi.SAMPLING_TILE):
            e.set_coords_2d(X0, Y0)
            st = e.run_2d(capi.layout_defaults(gd, iter_max=10, sampling=sampling))
            X, Y = e.get_coords_2d()
            assert np.all(np.isfinite(X)) and orc.path_stress_2d(go, X, Y, 200000, 1) < s0
            multi = int(np.repeat(counts > 1, counts).sum())
            if sampling == capi.SAMPLING_STREAM:
                assert st["term_updates"] == 10 * 10 * S            # 1-step paths are redrawn until the quota is met
            else:
                assert st["term_updates"] == 10 * 10 * multi        # tile sampling skips them (never counted)


def test_positions_beyond_32_bits():
    """Path positions above 2^32 bp: the 64-bit position field of the step record and the integer distance."""
    node_len = np.array([2_000_000_000, 1_900_000_000, 7, 2_100_000_000, 13, 1_500_000_000], dtype=np.uint32)
    first = np.array([0, 6, 11], dtype=np.uint64)
    step_node = np.array([0, 1, 2, 3, 4, 5, 5, 3, 1, 0, 2], dtype=np.uint32)
    step_rev = np.array([0, 1, 0, 0, 1, 0, 1, 1, 0, 0, 1], dtype=np.uint8)
    gd = odgi_b200.FlatGraph(node_len, first, step_node, step_rev)
    go = orc.Graph(node_len, first, step_node, step_rev)
    assert int(go.step_pos.max()) > 2 ** 32
    cd, co = capi.layout_defaults(gd), orc.default_layout_config(go)
    with odgi_b200.Engine(gd) as e:
        for cooling in (False, True):
            dev = e.sample_terms(cd, 2, cooling, 3000)
            ref, valid = orc.sample_terms(go, co, 2, cooling, 3000)
            ok = valid.astype(bool)
            for f in INT_FIELDS:
                assert np.array_equal(dev[f][ok], ref[f][ok].astype(dev[f].dtype)), f
    # single stream, strict order: arithmetic identical to the oracle's fp32 model at these magnitudes too
    kw = dict(iter_max=3, min_term_updates=500)
    cd, co = capi.layout_defaults(gd, n_streams=1, batch=1, sampling=capi.SAMPLING_STREAM, **kw), orc.default_layout_config(go, **kw)
    X0, Y0 = orc.layout_init(go, 1)
    X, Y, _ = odgi_b200.layout_2d(gd, cd, X0, Y0)
    _, xy = orc.layout_2d_f32(go, co, orc.XY_to_xy(X0, Y0), n_streams=1)
    Xr, Yr = orc.xy_to_XY(xy)
    assert np.array_equal(X, Xr) and np.array_equal(Y, Yr)


def test_scale_properties_on_a_graph_beyond_l2():
    """Size-independent properties at a size the oracle cannot run in seconds (6e5 nodes, 4.6e7 steps, 0.7 GB of step
    records): exact update counts, finite coordinates, the annealing half lowering the stress far below the initial
    one, tile and stream sampling converging to the same stress, device stress == oracle stress on the downloaded
    coordinates."""
    from odgi_b200 import synth
    gd = synth.preset("mid")
    go = orc.Graph(gd.node_len, gd.path_first_step, gd.step_node, gd.step_rev)
    X0, Y0 = odgi_b200.layout_init(gd, 42)
    finals = {}
    with odgi_b200.Engine(gd) as e:
        for sampling in (capi.SAMPLING_TILE, capi.SAMPLING_STREAM):
            cd = capi.layout_defaults(gd, sampling=sampling)
            e.set_coords_2d(X0, Y0)
            s_init = e.path_stress(2, 1_000_000, 3)
            mids = []
            for a, b in ((0, 15), (15, 22), (22, 30)):
                st = e.run_range(cd, 2, a, b)
                assert st["term_updates"] == (b - a) * 10 * gd.S and st["kernel_launches"] == b - a
                mids.append(e.path_stress(2, 1_000_000, 3))
            # the first half runs at learning rates that saturate every update; the annealing half must converge
            assert mids[2] < mids[1] and mids[2] < 0.05 * s_init, (s_init, mids)
            finals[sampling] = mids[2]
        X, Y = e.get_coords_2d()
    assert np.all(np.isfinite(X)) and np.all(np.isfinite(Y))
    assert finals[capi.SAMPLING_STREAM] == orc.path_stress_2d(go, X, Y, 1_000_000, 3)
    assert abs(finals[capi.SAMPLING_TILE] - finals[capi.SAMPLING_STREAM]) <= 0.05 * finals[capi.SAMPLING_STREAM], finals


def test_hub_node_switches_to_exchange_writes():
    """A node carrying a large share of all steps would see dozens of concurrent red.adds: the engine falls back to the
    reference kernel's last-writer-wins exchange by itself (stats.flags_used) and the layout stays finite."""
    rng = np.random.Generator(np.random.PCG64(11))
    N, P, L = 20_000, 8, 40_000
    node_len = rng.integers(1, 20, size=N).astype(np.uint32)
    first = (np.arange(P + 1) * L).astype(np.uint64)
    step_node = rng.integers(1, N, size=P * L).astype(np.uint32)
    step_node[::3] = 0                      # every third step of every path sits on node 0
    gd = odgi_b200.FlatGraph(node_len, first, step_node, np.zeros(P * L, dtype=np.uint8))
    go = orc.Graph(node_len, first, step_node, np.zeros(P * L, dtype=np.uint8))
    X0, Y0 = orc.layout_init(go, 1)
    X, Y, st = odgi_b200.layout_2d(gd, capi.layout_defaults(gd, iter_max=10), X0, Y0)
    assert st["flags_used"] & capi.PGSGD_FLAG_EXCH_WRITE
    assert np.all(np.isfinite(X)) and np.all(np.isfinite(Y))
    # an ordinary graph keeps the lossless default
    a = odgi_b200.FlatGraph(node_len, first, rng.integers(0, N, size=P * L).astype(np.uint32), np.zeros(P * L, dtype=np.uint8))
    _, _, st2 = odgi_b200.layout_2d(a, capi.layout_defaults(a, iter_max=3), X0, Y0)
    assert not (st2["flags_used"] & capi.PGSGD_FLAG_EXCH_WRITE)


def test_device_order_equals_reference_sort(graphs):
    """pgsgd_engine_order_1d == the (pos, handle) sort of path_linear_sgd_order on the same X — including ties"""
    gd, go = graphs["DRB1-3123"]
    with odgi_b200.Engine(gd) as e:
        e.set_coords_1d(None)
        e.run_1d(capi.sort_defaults(gd, iter_max=20))
        x = e.get_coords_1d()
        assert np.array_equal(e.order_1d(), orc.order_from_x(x))
        xt = np.round(x / 50.0) * 50.0 - 1e4       # many exact ties, negative values
        e.set_coords_1d(xt)
        assert np.array_equal(e.order_1d(), orc.order_from_x(xt))


def test_run_range_continues_the_schedule(graphs):
    """run_range(0, k) + run_range(k, n) == run(0, n): same streams, same schedule (single stream: bit-exact)"""
    gd, go = graphs["DRB1-3123"]
    cd = capi.layout_defaults(gd, iter_max=6, min_term_updates=3000, n_streams=1, batch=1, sampling=capi.SAMPLING_STREAM)
    X0, Y0 = orc.layout_init(go, 4)
    with odgi_b200.Engine(gd) as e:
        e.set_coords_2d(X0, Y0)
        e.run_2d(cd)
        full = e.get_coords_2d_f32()
        e.set_coords_2d(X0, Y0)
        a = e.run_range(cd, 2, 0, 2)
        b = e.run_range(cd, 2, 2, 6)
        split = e.get_coords_2d_f32()
    assert a["iterations_run"] == 2 and b["iterations_run"] == 4
    assert np.array_equal(full, split)


def test_degenerate_graphs_do_nothing_like_the_reference():
    """No path with more than one step: the reference returns without touching the layout (path_sgd_layout.cpp:64-74;
    its own unit test of this corner: src/unittest/sort.cpp:129-255 — ten 1-node paths).  Also: a graph without paths."""
    node_len = np.array([1] * 10, dtype=np.uint32)
    first = np.arange(11, dtype=np.uint64)                      # ten paths of one step each
    step_node = np.arange(10, dtype=np.uint32)
    gd = odgi_b200.FlatGraph(node_len, first, step_node, np.zeros(10, dtype=np.uint8))
    X0 = np.arange(20, dtype=np.float64)
    Y0 = np.arange(20, dtype=np.float64) * 0.5
    X, Y, st = odgi_b200.layout_2d(gd, capi.layout_defaults(gd, min_term_updates=1000, eta_max=4.0, space=1), X0, Y0)
    assert st["term_updates"] == 0 and st["iterations_run"] == 0
    assert np.array_equal(X, X0) and np.array_equal(Y, Y0)
    x, st = odgi_b200.sort_1d(gd, capi.sort_defaults(gd, min_term_updates=1000, eta_max=4.0, space=1))
    assert st["term_updates"] == 0 and np.array_equal(x, np.arange(10, dtype=np.float64))   # cumulative-bp initialisation
    empty = odgi_b200.FlatGraph(node_len, np.zeros(1, dtype=np.uint64), np.zeros(0, dtype=np.uint32), np.zeros(0, dtype=np.uint8))
    X, Y, st = odgi_b200.layout_2d(empty, capi.Config(iter_max=3, min_term_updates=10, eta_max=4.0, space=1), X0, Y0)
    assert st["term_updates"] == 0 and np.array_equal(X, X0)
