"""`.lay` encoded on the device (pgsgd_engine_encode_lay) == the host writer's bytes for the same coordinates.
The host writer (odgi_b200/host/lay_format.hpp, `pgsgd lay -c`) is pinned byte-for-byte on files the reference's own
Layout::serialize wrote (tests/test_host_cpu.py), so equality here is equality with src/algorithms/layout.cpp:43-61 — including
the per-component stacking of src/subcommand/layout_main.cpp:402-435 when component ids are given."""
import os
import subprocess

import numpy as np
import pytest

import odgi_b200
from odgi_b200 import capi
from odgi_b200.arrays import read_arrays, write_arrays
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "odgi_b200", "host", "pgsgd")


def host_lay(tmp_path, X, Y, tag):
    plain, out = tmp_path / f"{tag}.arr", tmp_path / f"{tag}.lay"
    write_arrays(str(plain), {"X": X, "Y": Y})
    subprocess.run([CLI, "lay", "-c", str(plain), "-o", str(out)], check=True)
    return out.read_bytes()


def stack_components(X, Y, comp):
    """layout_main.cpp:406-433 in numpy (fp64)"""
    X, Y = X.copy(), Y.copy()
    border, curr = 1000.0, 1000.0
    for k in range(int(comp.max()) + 1):
        idx = np.nonzero(comp == k)[0]
        j = np.concatenate([2 * idx, 2 * idx + 1])
        min_x, min_y, max_y = X[j].min(), Y[j].min(), Y[j].max()
        x_off, y_off = min_x - border, curr - min_y
        curr += (max_y - min_y) + border
        X[j] -= x_off
        Y[j] += y_off
    return X, Y


@pytest.mark.parametrize("name", ["DRB1-3123", "k", "note5"])
def test_device_lay_is_byte_identical_to_the_host_writer(tmp_path, name):
    a = read_arrays(os.path.join(GOLDEN, f"{name}.graph.arr.gz"))
    gd, go = odgi_b200.graph_from_arrays(a), orc.Graph.from_arrays(a)
    X0, Y0 = orc.layout_init(go, seed=7)
    with odgi_b200.Engine(gd) as e:
        e.set_coords_2d(X0, Y0)
        e.run_range(capi.layout_defaults(gd, iter_max=6), 2, 0, 6)
        X, Y = e.get_coords_2d()
        assert e.encode_lay() == host_lay(tmp_path, X, Y, "plain")
        # three artificial components (ids interleaved): the stacking runs on the device too
        comp = (np.arange(gd.N) % 3).astype(np.uint32)
        Xs, Ys = stack_components(X, Y, comp)
        assert e.encode_lay(comp) == host_lay(tmp_path, Xs, Ys, "stacked")
        # one component: still moved to the (border, border) corner
        one = np.zeros(gd.N, dtype=np.uint32)
        Xs, Ys = stack_components(X, Y, one)
        assert e.encode_lay(one) == host_lay(tmp_path, Xs, Ys, "one")


def test_device_lay_degenerate_values(tmp_path):
    """equal neighbours (delta 0 is coded as 2^64), negative and huge coordinates, a value count that is not a multiple of 128"""
    n = 77
    lens = np.ones(n, dtype=np.uint32)
    g = capi.FlatGraph(lens, np.array([0, n], dtype=np.uint64), np.arange(n, dtype=np.uint32), None, None)
    rng = np.random.default_rng(3)
    X = np.repeat(rng.normal(0, 1e7, n).astype(np.float32).astype(np.float64), 2)     # both ends equal -> zero deltas
    Y = np.zeros(2 * n)
    Y[::5] = -3.5e8
    with odgi_b200.Engine(g) as e:
        e.set_coords_2d(X, Y)
        Xd, Yd = e.get_coords_2d()
        assert e.encode_lay() == host_lay(tmp_path, Xd, Yd, "deg")
