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


def assign_paths(step_counts, n_ranks: int):
    """Deal the paths of a job out over n_ranks: greedy bin packing on step count, longest path first, ties
    to the lower rank.  Returns the owner rank of every path.  A term always pairs two steps of the same path, so whole
    paths are the unit; the ranks' shares of an iteration's updates follow their step counts."""
    step_counts = np.asarray(step_counts, dtype=np.uint64)
    owner = np.zeros(step_counts.size, dtype=np.int64)
    load = np.zeros(n_ranks, dtype=np.uint64)
    for p in np.argsort(-step_counts.astype(np.int64), kind="stable"):
        r = int(np.argmin(load))
        owner[p] = r
        load[r] += step_counts[p]
    return owner


def shard_paths(g: FlatGraph, n_ranks: int, rank: int) -> FlatGraph:
    """The view rank `rank` creates its engine from: the whole node table, only the steps of ITS paths (in path order).
    Use with Engine.set_shard(g.S)."""
    first = g.path_first_step
    owner = assign_paths(np.diff(first), n_ranks)
    mine = np.nonzero(owner == rank)[0]
    sel = np.concatenate([np.arange(first[p], first[p + 1], dtype=np.int64) for p in mine]) if mine.size else np.zeros(0, dtype=np.int64)
    new_first = np.concatenate([[0], np.cumsum(np.diff(first)[mine])]).astype(np.uint64)
    names = [g.path_names[p] for p in mine] if g.path_names else []
    return FlatGraph(g.node_len, new_first, g.step_node[sel], None if g.step_rev is None else g.step_rev[sel],
                     None if g.step_pos is None else g.step_pos[sel], names)
