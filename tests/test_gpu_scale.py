"""Parity AT SCALE, anchored to the CPU reference: default runs (AUTO sampling: the tile kernel bench.py
times on `mid`, the stream kernel on the shallow `longthin`) against full default runs of the UNMODIFIED reference CPU implementation on the same graphs from the
same injected initialisation (tests/golden/stress_reference_scale.json, made by scripts/make_scale_golden.py; where the final
stress depends on the worker-stream seeds more than on thread timing — the reference hard-codes its seeds — the band also holds
runs of the oracle, the bit-exact restatement of the reference, under other seeds: DESIGN.md 5.4):

  mid       6.0e5 nodes, 4.6e7 steps (90 haplotypes)
  longthin  3.6e6 nodes, 1.8e7 steps (6 haplotypes), path length 4e7 bp: layout coordinates beyond 2^24, where an fp32
            coordinate no longer resolves a base pair (ulp 4 at 4e7) — the reference computes in fp64

Two readouts, both evaluated on the device with the oracle's definitions (bit-identical sums, checked below on a small graph):
the far-pair sampled path stress (partner uniform in the path) and the LOCAL stress (partner 1..64 ranks away, <= 1000 bp), which
is the one that would show lost coordinate precision.  Criterion: SURVEY.md 8(d), see _assert_in_band."""
import json
import os

import numpy as np
import pytest

import odgi_b200
from odgi_b200 import capi, synth
from odgi_b200.arrays import read_arrays
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bands():
    p = os.path.join(GOLDEN, "stress_reference_scale.json")
    if not os.path.exists(p):
        return {}
    with open(p) as f:
        return json.load(f)


def _assert_in_band(values, band, what, upper_only=False):
    """SURVEY.md 8(d) on the mean over the seeds, max(1 %, 3 sigma) per run (see tests/test_gpu_parity.py).
    upper_only (the LOCAL stress): fp32 coordinates of magnitude M are quantised to M * 2^-23 (4 bp at 4e7), so node ends closer
    than that collapse onto one point: the fp32 format can only LOWER the local stress (oracle fp32 model: -2 % at 3e7, -9 % at
    1.2e8).  The gate is then "not above the band", plus a floor at half the band mean against a degenerate layout."""
    mean, sd = band["mean"], band["sd"]
    tol, tol1 = max(0.01 * mean, 2 * sd), max(0.01 * mean, 3 * sd)
    if upper_only:
        assert 0.5 * mean <= np.mean(values) <= mean + tol, (what, values, mean, sd)
        assert all(v <= mean + tol1 for v in values), (what, values, mean, sd)
    else:
        assert abs(np.mean(values) - mean) <= tol, (what, values, mean, sd)
        assert all(abs(v - mean) <= tol1 for v in values), (what, values, mean, sd)


def test_device_local_stress_equals_the_oracle():
    """pgsgd_engine_local_stress == orc_local_stress_* (same generator streams, IEEE fp64 sums: bit-identical), 2D and 1D"""
    a = read_arrays(os.path.join(GOLDEN, "DRB1-3123.graph.arr.gz"))
    gd, go = odgi_b200.graph_from_arrays(a), orc.Graph.from_arrays(a)
    X0, Y0 = orc.layout_init(go, seed=3)
    with odgi_b200.Engine(gd) as e:
        e.set_coords_2d(X0, Y0)
        e.run_range(capi.layout_defaults(gd, iter_max=4), 2, 0, 4)
        X, Y = e.get_coords_2d()
        assert e.local_stress(2, 300_000, 77) == orc.local_stress_2d(go, X, Y, 300_000, 77)
        assert e.path_stress(2, 300_000, 77) == orc.path_stress_2d(go, X, Y, 300_000, 77)
        x = orc.sort_init(go) * 1.0009765625
        e.set_coords_1d(x)
        assert e.local_stress(1, 300_000, 78) == orc.local_stress_1d(go, x, 300_000, 78)


SEEDS_2D = [9399220, 1234567, 42] + [1000003 * (k + 1) for k in range(13)]   # far apart: worker stream t of a run is seeded seed + t


def _assert_mean_in_band(values, band, what, upper_only=False):
    """SURVEY.md 8(d) on the mean — within max(1 %, 2 sigma) of the band — with the standard error of OUR mean added to the tolerance:
    on these graphs the final stress of a GPU run scatters 2-3x wider over seeds than the band's runs do (the reference's own CUDA
    path scatters by as much from GPU timing alone, `refcuda` in the golden file), so a mean over n seeds is only known to
    2 * sd / sqrt(n).  Without that term the outcome would depend on which seeds the test happens to use."""
    mean, sd = band["mean"], band["sd"]
    v = np.asarray(values)
    tol = max(0.01 * mean, 2 * sd) + 2 * v.std(ddof=1) / np.sqrt(len(v))
    if upper_only:
        assert 0.5 * mean <= v.mean() <= mean + tol, (what, v.mean(), v.std(ddof=1), mean, sd)
    else:
        assert abs(v.mean() - mean) <= tol, (what, v.mean(), v.std(ddof=1), mean, sd)


@pytest.mark.parametrize("name", ["longthin", "mid"])
def test_default_2d_run_within_the_reference_band_at_scale(name):
    """Two anchors: the CPU reference band (reference runs + oracle runs under other seeds) and the band of the
    reference's OWN CUDA path on a B200 (`refcuda`: the implementation this library drops in for).  16 seeds, because single runs
    scatter widely (sd 20-30 % of the mean, for this library's two samplers and for the reference CUDA path alike)."""
    band = _bands().get(f"{name}.layout2d")
    if band is None or band["runs"] < 3:
        pytest.skip(f"no reference band (>= 3 runs) for {name}.layout2d yet (scripts/make_scale_golden.py)")
    g = synth.generate(*band["generator"], seed=42)
    assert g.S == band["steps"] and g.N == band["nodes"] and g.S >= (1 << 22)
    X0, Y0 = odgi_b200.layout_init(g, seed=band["init_seed"])
    far, loc = [], []
    with odgi_b200.Engine(g) as e:
        for seed in SEEDS_2D:
            cd = capi.layout_defaults(g, seed=seed)
            e.set_coords_2d(X0, Y0)
            st = e.run_2d(cd)
            assert st["term_updates"] == 30 * 10 * g.S and not (st["flags_used"] & capi.FLAG_LEGACY_TILE)
            # AUTO sampling: tile for mid (76 steps per node), stream for longthin (5 steps per node: include/pgsgd.h)
            assert st["sampling_used"] == (capi.SAMPLING_TILE if name == "mid" else capi.SAMPLING_STREAM)
            far.append(e.path_stress(2, band["n_pairs"], band["seed"]))
            loc.append(e.local_stress(2, band["n_pairs"], band["seed"]))
        X, _ = e.get_coords_2d()
    if name == "longthin":
        assert np.max(np.abs(X)) > 2 ** 24   # the regime this graph is here for
    # (1) against the CPU reference: the mean, two-sided for the far stress, from above for the local stress (fp32, see _assert_in_band)
    _assert_mean_in_band(far, band["far"], f"{name} far vs the CPU reference")
    _assert_mean_in_band(loc, band["local"], f"{name} local vs the CPU reference", upper_only=True)
    # (2) against the reference's own CUDA path: not worse than it — means not above its band, no run's far stress beyond its mean + 3 sigma
    rc = band.get("refcuda")
    if rc is not None:
        for vals, b, what in ((far, rc["far"], "far"), (loc, rc["local"], "local")):
            assert np.mean(vals) <= b["mean"] + max(0.01 * b["mean"], 2 * b["sd"]), (name, what, np.mean(vals), b["mean"], b["sd"])
        assert max(far) <= rc["far"]["mean"] + max(0.01 * rc["far"]["mean"], 3 * rc["far"]["sd"]), (name, max(far), rc["far"])


@pytest.mark.parametrize("name", ["longthin", "mid"])
def test_default_1d_within_the_reference_band_at_scale(name):
    band = _bands().get(f"{name}.sort1d")
    if band is None or band["runs"] < 3:
        pytest.skip(f"no reference band (>= 3 runs) for {name}.sort1d yet (scripts/make_scale_golden.py)")
    g = synth.generate(*band["generator"], seed=42)
    far, loc = [], []
    with odgi_b200.Engine(g) as e:
        for seed in (9399220, 1234567, 42):
            cd = capi.sort_defaults(g, seed=seed)
            e.set_coords_1d(None)
            st = e.run_1d(cd)
            assert st["iterations_run"] == 101
            far.append(e.path_stress(1, band["n_pairs"], band["seed"]))
            loc.append(e.local_stress(1, band["n_pairs"], band["seed"]))
    _assert_in_band(far, band["far"], f"{name} 1D far")
    _assert_in_band(loc, band["local"], f"{name} 1D local", upper_only=True)
