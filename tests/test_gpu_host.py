"""The host side end to end on a GPU: the standalone CLI and (when oracle/_ref was built) the reference's own call
chain with our shim linked in place of src/cuda/layout.cu."""
import json
import os
import subprocess

import numpy as np
import pytest

import odgi_b200
from odgi_b200 import synth
from odgi_b200.arrays import read_arrays
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "odgi_b200", "host", "pgsgd")
SHIM = os.path.join(ROOT, "oracle", "_ref", "shim_driver")


def _band(name):
    with open(os.path.join(ROOT, "tests", "golden", "stress_reference.json")) as f:
        return json.load(f)[name]


@pytest.fixture(scope="module")
def drb1(golden_graphs, tmp_path_factory):
    a = golden_graphs["DRB1-3123"]
    g = odgi_b200.graph_from_arrays(a)
    gfa = tmp_path_factory.mktemp("gfa") / "DRB1-3123.gfa"
    synth.write_gfa(g, str(gfa))
    return str(gfa), orc.Graph.from_arrays(a)


def test_cli_layout_device_ingest(drb1, tmp_path):
    """`pgsgd layout --device-ingest`: GFA lines found on the host, step lists parsed / flattened / laid out / stacked / .lay-encoded
    on the GPU.  The .lay decodes to a layout in the reference's stress band; the TSV route gives the same component column."""
    gfa, go = drb1
    lay, back, tsv = tmp_path / "dev.lay", tmp_path / "dev.arr", tmp_path / "dev.tsv"
    r = subprocess.run([CLI, "layout", "-i", gfa, "-o", str(lay), "-T", str(tsv), "--gpu", "--init-seed", "42", "--device-ingest", "--timing"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads([ln for ln in r.stderr.splitlines() if ln.startswith("{")][-1])
    assert info["ingest"] == "device" and info["nodes"] == go.N and info["steps"] == go.S and info["term_updates"] == 30 * 10 * go.S
    subprocess.run([CLI, "lay", "-i", str(lay), "-a", str(back)], check=True)
    b = read_arrays(str(back))
    assert b["X"].size == 2 * go.N and np.all(np.isfinite(b["X"])) and np.all(np.isfinite(b["Y"]))
    band = _band("DRB1-3123.layout2d")
    s = orc.path_stress_2d(go, b["X"], b["Y"], band["n_pairs"], band["seed"])   # the stacking is a translation: stress unchanged
    assert abs(s - band["mean"]) <= 0.03 * band["mean"], (s, band["mean"])
    rows = np.loadtxt(str(tsv), skiprows=1)
    assert rows.shape == (2 * go.N, 4) and np.allclose(np.sort(rows[:, 0]), np.arange(2 * go.N))
    # one component, moved to the (border, border) corner by layout_main.cpp:402-433
    assert abs(b["X"].min() - 1000.0) < 1e-6 and abs(b["Y"].min() - 1000.0) < 1e-6


def test_cli_layout_tsv(drb1, tmp_path):
    gfa, go = drb1
    tsv = tmp_path / "lay.tsv"
    subprocess.run([CLI, "layout", "-i", gfa, "-T", str(tsv), "--gpu", "--init-seed", "42", "-P"], check=True)
    rows = np.loadtxt(str(tsv), skiprows=1)
    assert rows.shape == (2 * go.N, 4) and np.array_equal(rows[:, 0], np.arange(2 * go.N))
    band = _band("DRB1-3123.layout2d")
    s = orc.path_stress_2d(go, rows[:, 1], rows[:, 2], band["n_pairs"], band["seed"])
    assert abs(s - band["mean"]) <= 0.03 * band["mean"], (s, band["mean"])


def test_cli_sort_order(drb1, tmp_path):
    gfa, go = drb1
    out, lay = tmp_path / "order.txt", tmp_path / "lay1d.tsv"
    subprocess.run([CLI, "sort", "-i", gfa, "-o", str(out), "-Y", "--gpu", "--layout-out", str(lay)], check=True)
    order = np.loadtxt(str(out), dtype=np.int64)
    assert np.array_equal(np.sort(order), np.arange(1, go.N + 1))
    rows = np.loadtxt(str(lay), skiprows=1)
    x = np.empty(go.N)
    x[rows[:, 0].astype(np.int64) - 1] = rows[:, 1]
    assert np.array_equal(order - 1, orc.order_from_x(x).astype(np.int64))  # the reference's (pos, handle) sort on the same X
    band = _band("DRB1-3123.sort1d")
    s = orc.path_stress_1d(go, x, band["n_pairs"], band["seed"])
    assert abs(s - band["mean"]) <= 0.03 * band["mean"] + 2 * band["sd"], (s, band["mean"])


@pytest.mark.skipif(not os.path.exists(SHIM), reason="oracle/_ref/shim_driver not built (needs the reference tree at build time)")
def test_reference_call_chain_with_our_shim(drb1, tmp_path):
    """algorithms::path_linear_sgd_layout_gpu (UNMODIFIED reference code) -> cuda::gpu_layout (our shim) -> C-ABI."""
    gfa, go = drb1
    out = tmp_path / "shim.arr"
    subprocess.run([SHIM, "layout", gfa, str(out)], check=True, cwd=str(tmp_path))
    r = read_arrays(str(out))
    band = _band("DRB1-3123.layout2d")
    s0 = orc.path_stress_2d(go, r["X0"], r["Y0"], band["n_pairs"], band["seed"])
    s = orc.path_stress_2d(go, r["X"], r["Y"], band["n_pairs"], band["seed"])
    assert s0 > 5 and abs(s - band["mean"]) <= 0.03 * band["mean"], (s0, s, band["mean"])
    out1 = tmp_path / "shim1d.arr"
    subprocess.run([SHIM, "sort", gfa, str(out1)], check=True, cwd=str(tmp_path))
    x = read_arrays(str(out1))["X"]
    b1 = _band("DRB1-3123.sort1d")
    s1 = orc.path_stress_1d(go, x, b1["n_pairs"], b1["seed"])
    assert abs(s1 - b1["mean"]) <= 0.03 * b1["mean"] + 2 * b1["sd"], (s1, b1["mean"])


@pytest.mark.skipif(not os.path.exists(SHIM) or odgi_b200.device_count() < 2, reason="needs oracle/_ref/shim_driver and 2 GPUs")
def test_reference_call_chain_on_two_gpus(drb1, tmp_path):
    """The same unmodified reference call chain, one process, PGSGD_GPUS=2: the one-shot C-ABI call fans out over host threads."""
    gfa, go = drb1
    env = dict(os.environ, PGSGD_GPUS="2")
    out = tmp_path / "shim2.arr"
    subprocess.run([SHIM, "layout", gfa, str(out)], check=True, cwd=str(tmp_path), env=env, timeout=300)
    r = read_arrays(str(out))
    band = _band("DRB1-3123.layout2d")
    s = orc.path_stress_2d(go, r["X"], r["Y"], band["n_pairs"], band["seed"])
    assert abs(s - band["mean"]) <= 0.03 * band["mean"], (s, band["mean"])
    out1 = tmp_path / "shim2_1d.arr"
    subprocess.run([SHIM, "sort", gfa, str(out1)], check=True, cwd=str(tmp_path), env=env, timeout=300)
    x = read_arrays(str(out1))["X"]
    b1 = _band("DRB1-3123.sort1d")
    s1 = orc.path_stress_1d(go, x, b1["n_pairs"], b1["seed"])
    assert abs(s1 - b1["mean"]) <= 0.03 * b1["mean"] + 2 * b1["sd"], (s1, b1["mean"])
