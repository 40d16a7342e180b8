"""path_linear_sgd_order's sort (src/algorithms/path_sgd.cpp:552-658) pinned on reference output: X and the
order of ONE reference run (tests/golden/order_*.json, scripts/make_order_golden.py).  multi3 has three weak components
with interleaved node ids, so the component key of the sort is observable."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    with open(os.path.join(GOLDEN, f"order_{name}.json")) as f:
        d = json.load(f)
    x = np.array([float.fromhex(v) for v in d["X"]], dtype=np.float64)
    order = np.array(d["order"], dtype=np.uint64)
    return d, x, order


def parse_small_gfa(text):
    """node lengths, per-path node ranks, and L edges of a small GFA (ids 1..N)"""
    node_len, paths, edges = {}, [], []
    for line in text.splitlines():
        f = line.split("\t")
        if f[0] == "S":
            node_len[int(f[1])] = len(f[2])
        elif f[0] == "P":
            paths.append([int(s[:-1]) - 1 for s in f[2].split(",")])
        elif f[0] == "L":
            edges.append((int(f[1]) - 1, int(f[3]) - 1))
    n = len(node_len)
    first = np.cumsum([0] + [len(p) for p in paths]).astype(np.uint64)
    step_node = np.array([v for p in paths for v in p], dtype=np.uint32)
    return n, first, step_node, edges


def multi3():
    d, x, order = load_fixture("multi3")
    n, first, step_node, edges = parse_small_gfa(d["gfa"])
    comp = orc.component_keys(n, first, step_node, edges)
    return x, order, comp


def test_reference_order_keys_on_the_weak_component():
    x, order, comp = multi3()
    assert len(set(comp.tolist())) == 3
    assert np.all(order & np.uint64(1) == 0)   # forward handles
    ref_ranks = order >> np.uint64(1)
    assert np.array_equal(orc.order_from_x(x, comp), ref_ranks)
    # the fixture discriminates: without the component key the order is a different one
    assert not np.array_equal(orc.order_from_x(x), ref_ranks)


def test_reference_order_single_component():
    _, x, order = load_fixture("DRB1-3123")
    assert np.array_equal(orc.order_from_x(x), order >> np.uint64(1))


@pytest.mark.gpu
def test_device_order_equals_the_reference_order():
    """pgsgd_engine_order_1d[_components] on the reference's X == the reference's order (one and three components)"""
    import odgi_b200
    from odgi_b200.arrays import read_arrays
    from odgi_b200 import capi
    # three components: the engine only needs N, so any path set over 40 nodes will do — use the fixture's own
    d, x, order = load_fixture("multi3")
    n, first, step_node, edges = parse_small_gfa(d["gfa"])
    lens = np.ones(n, dtype=np.uint32)
    g = capi.FlatGraph(lens, first, step_node, None, None)
    comp = orc.component_keys(n, first, step_node, edges)
    with odgi_b200.Engine(g) as e:
        e.set_coords_1d(x)
        assert np.array_equal(e.order_1d(comp), order >> np.uint64(1))
        assert np.array_equal(e.order_1d(), orc.order_from_x(x))
    _, x, order = load_fixture("DRB1-3123")
    gd = odgi_b200.graph_from_arrays(read_arrays(os.path.join(GOLDEN, "DRB1-3123.graph.arr.gz")))
    with odgi_b200.Engine(gd) as e:
        e.set_coords_1d(x)
        assert np.array_equal(e.order_1d(), order >> np.uint64(1))
