"""CPU checks of the stress bands the GPU parity tests are gated on (tests/golden/stress_reference*.json): that they hold what the
tests read, that every band rests on enough runs, and that the oracle reproduces a stored oracle-seed value (the band files are
reproducible from scripts/make_stress_golden.py --oracle-seeds and scripts/make_scale_golden.py, not hand-edited)."""
import json
import os

import numpy as np

from odgi_b200.arrays import read_arrays
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def test_small_graph_bands_hold_reference_and_oracle_seed_runs():
    bands = _load("stress_reference.json")
    for key in ("DRB1-3123.layout2d", "chr6.C4.layout2d", "LPA.layout2d", "DRB1-3123.sort1d", "chr6.C4.sort1d", "LPA.sort1d"):
        b = bands[key]
        ref, orc_runs = b["reference_only"]["values"], b["oracle_seed_runs"]["values"]
        assert len(ref) >= 6 and len(orc_runs) >= 12 and len(set(b["oracle_seed_runs"]["seeds"])) == len(orc_runs)
        allv = np.array(list(ref) + list(orc_runs))
        assert np.isclose(b["mean"], allv.mean(), rtol=1e-12) and np.isclose(b["sd"], allv.std(ddof=1), rtol=1e-12)
        assert b["n_pairs"] >= 100_000 and b["mean"] > 0
        if key.endswith("layout2d"):
            assert b["mean"] < b["initial"]   # (a 1D sort may end above its initial order's stress: chr6.C4, reference and oracle alike)


def test_scale_bands_hold_reference_oracle_and_reference_cuda_runs():
    bands = _load("stress_reference_scale.json")
    for key in ("mid.layout2d", "longthin.layout2d", "mid.sort1d", "longthin.sort1d"):
        b = bands[key]
        assert b["runs"] >= 3 and b["runs"] == b["reference_runs"] + len(b["oracle_runs"]) == len(b["far"]["values"]) == len(b["local"]["values"])
        for m in ("far", "local"):
            v = np.array(b[m]["values"])
            assert np.isclose(b[m]["mean"], v.mean(), rtol=1e-12) and np.isclose(b[m]["sd"], v.std(ddof=1), rtol=1e-12)
        assert b["far"]["mean"] < b["initial_far"]
    for key in ("mid.layout2d", "longthin.layout2d"):   # the second anchor of tests/test_gpu_scale.py
        rc = bands[key]["refcuda"]
        assert rc["runs"] >= 8 and len(rc["far"]["values"]) == rc["runs"] and rc["far"]["sd"] > 0
    # the 2D bands carry oracle runs under seeds other than the reference's hard-coded ones
    assert all(len(bands[k]["oracle_runs"]) >= 5 for k in ("mid.layout2d", "longthin.layout2d", "mid.sort1d", "longthin.sort1d"))


def test_recorded_gpu_1d_values_sit_inside_the_1d_scale_bands():
    """profiles/r02_1d_scale_values.jsonl: what the default 1D runs of tests/test_gpu_scale.py produced on a B200 (same seeds; a run
    repeats to ~1 %).  Evaluated here with that test's own criterion, so that a band update is checked without a GPU."""
    rec_path = os.path.join(os.path.dirname(GOLDEN), "..", "profiles", "r02_1d_scale_values.jsonl")
    if not os.path.exists(rec_path):
        import pytest
        pytest.skip("no recorded GPU values")
    bands = _load("stress_reference_scale.json")
    for ln in open(rec_path):
        r = json.loads(ln)
        b = bands[f"{r['graph']}.sort1d"]
        for m, upper in (("far", False), ("local", True)):
            mean, sd, v = b[m]["mean"], b[m]["sd"], np.array(r[m])
            tol, tol1 = max(0.01 * mean, 2 * sd), max(0.01 * mean, 3 * sd)
            if upper:
                assert 0.5 * mean <= v.mean() <= mean + tol and all(x <= mean + tol1 for x in v), (r["graph"], m, v, mean, sd)
            else:
                assert abs(v.mean() - mean) <= tol and all(abs(x - mean) <= tol1 for x in v), (r["graph"], m, v, mean, sd)


def test_a_stored_oracle_seed_run_is_reproducible():
    """DRB1-3123 2D, oracle seed 42, six interleaved streams: the stored value is what the oracle computes today"""
    b = _load("stress_reference.json")["DRB1-3123.layout2d"]
    i = b["oracle_seed_runs"]["seeds"].index(42)
    g = orc.Graph.from_arrays(read_arrays(os.path.join(GOLDEN, "DRB1-3123.graph.arr.gz")))
    cfg = orc.default_layout_config(g)
    cfg.seed = 42
    X, Y = orc.layout_init(g, seed=b.get("init_seed", 42))
    _, X, Y = orc.layout_2d(g, cfg, X, Y, n_streams=b["oracle_seed_runs"]["n_streams"])
    assert orc.path_stress_2d(g, X, Y, b["n_pairs"], b["seed"]) == b["oracle_seed_runs"]["values"][i]


def test_recorded_gpu_2d_values_sit_inside_the_2d_scale_bands():
    """profiles/r02_2d_scale_values.jsonl: the 16-seed default 2D runs of tests/test_gpu_scale.py as recorded on a B200, evaluated with
    that test's criterion (mean against the CPU band with the standard error of our mean; not above the reference CUDA path's band)."""
    rec_path = os.path.join(os.path.dirname(GOLDEN), "..", "profiles", "r02_2d_scale_values.jsonl")
    if not os.path.exists(rec_path):
        import pytest
        pytest.skip("no recorded GPU values")
    bands = _load("stress_reference_scale.json")
    for ln in open(rec_path):
        r = json.loads(ln)
        b = bands[f"{r['graph']}.layout2d"]
        for m, upper in (("far", False), ("local", True)):
            mean, sd, v = b[m]["mean"], b[m]["sd"], np.array(r[m], dtype=np.float64)
            tol = max(0.01 * mean, 2 * sd) + 2 * v.std(ddof=1) / np.sqrt(len(v))
            if upper:
                assert 0.5 * mean <= v.mean() <= mean + tol, (r["graph"], m, v.mean(), mean, sd)
            else:
                assert abs(v.mean() - mean) <= tol, (r["graph"], m, v.mean(), mean, sd)
            rc = b["refcuda"][m]
            assert v.mean() <= rc["mean"] + max(0.01 * rc["mean"], 2 * rc["sd"]), (r["graph"], m, v.mean(), rc["mean"], rc["sd"])
        rc = b["refcuda"]["far"]
        assert max(r["far"]) <= rc["mean"] + max(0.01 * rc["mean"], 3 * rc["sd"]), (r["graph"], max(r["far"]), rc)
