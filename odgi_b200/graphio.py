"""Flattened-graph I/O: PGSGDARR containers <-> FlatGraph, and initial coordinates as `odgi layout` makes them."""
from __future__ import annotations

import numpy as np

from .arrays import read_arrays, write_arrays
from .capi import FlatGraph


def graph_from_arrays(a) -> FlatGraph:
    names = bytes(a["path_names"]).decode().split("\n")[:-1] if "path_names" in a else []
    return FlatGraph(a["node_len"], a["path_first_step"], a["step_node"], a.get("step_rev"), a.get("step_pos"), names)


def load_graph_arrays(path: str) -> FlatGraph:
    return graph_from_arrays(read_arrays(path))


def save_graph_arrays(path: str, g: FlatGraph) -> None:
    arrs = {"node_len": g.node_len, "path_first_step": g.path_first_step, "step_node": g.step_node}
    if g.step_rev is not None:
        arrs["step_rev"] = g.step_rev
    if g.step_pos is not None:
        arrs["step_pos"] = g.step_pos
    write_arrays(path, arrs)


def layout_init(g: FlatGraph, seed: int = 42, noise: bool = True):
    """`odgi layout` default ('d') initialisation (layout_main.cpp:322-328): X = cumulative bp of the node
    order for both node ends, Y ~ N(0, sqrt(2N)).  The reference seeds from std::random_device; a fixed
    seed is injected here so runs are reproducible."""
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
