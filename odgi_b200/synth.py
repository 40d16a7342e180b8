"""Synthetic pangenome graphs in flattened form (SURVEY.md §8d generator spec).

A backbone of `n_sites` sites; every `bubble_every`-th site is a bubble with 2-4 alleles (one node per
allele), the others have one node.  Node length is 1 bp w.p. 0.55, else 1 + Geom(mean 24) capped at 1024
(mean ~ 12 bp, SNP-dense like PGGB graphs).  Each of `n_paths` haplotypes walks all sites choosing alleles
i.i.d. with per-site frequencies ~ Dirichlet-ish; a few segments per path are inverted (reverse-orientation
steps) or tandem-duplicated x2-5 (cycles).  Node ids are compact 1..N by construction.
"""
from __future__ import annotations

import numpy as np

from .capi import FlatGraph

PRESETS = {
    # name: (n_sites, n_paths)        nodes ~ 1.2 * n_sites, steps ~ 1.01 * n_sites * n_paths
    "tiny": (2_000, 8),
    "small": (50_000, 16),
    "mid": (500_000, 90),             # ~6e5 nodes, ~4.5e7 steps  (0.7 GB of step records: beyond L2)
    "c4": (4_600_000, 90),            # ~5.5e6 nodes, ~4.2e8 steps (BASELINE config 4: 90-haplotype chr6-MHC scale)
    "c4x": (5_500_000, 90),           # ~6.6e6 nodes, ~5.0e8 steps (the same generator at the upper reading of "~5e8 path steps")
}


def generate(n_sites: int, n_paths: int, seed: int = 42, bubble_every: int = 10, inv_per_mbp: float = 0.5,
             dup_per_mbp: float = 0.2, with_pos: bool = False) -> FlatGraph:
    rng = np.random.Generator(np.random.PCG64(seed))
    # alleles per site
    k = np.ones(n_sites, dtype=np.int64)
    bub = np.arange(bubble_every // 2, n_sites, bubble_every)
    k[bub] = rng.integers(2, 5, size=bub.size)
    site_first = np.zeros(n_sites + 1, dtype=np.int64)
    np.cumsum(k, out=site_first[1:])
    N = int(site_first[-1])
    # node lengths
    node_len = np.ones(N, dtype=np.uint32)
    longer = rng.random(N) >= 0.55
    node_len[longer] = 1 + np.minimum(rng.geometric(1.0 / 24.0, size=int(longer.sum())), 1023).astype(np.uint32)
    # per-bubble allele frequency thresholds (U-shaped minor allele frequencies)
    maf = rng.beta(0.5, 0.5, size=bub.size) * 0.5

    mean_len = float(node_len.mean())
    site_bp = 1e6 / mean_len  # sites per Mbp
    step_node_parts, step_rev_parts, counts = [], [], []
    for _ in range(n_paths):
        allele = np.zeros(n_sites, dtype=np.int64)
        u = rng.random(bub.size)
        # minor alleles share `maf` evenly; allele 0 is the major one
        minor = u < maf
        kk = k[bub] - 1
        allele[bub] = np.where(minor, 1 + np.minimum((u / np.maximum(maf, 1e-12) * kk).astype(np.int64), kk - 1), 0)
        nodes = (site_first[:-1] + allele).astype(np.uint32)
        rev = np.zeros(n_sites, dtype=np.uint8)
        # structural events: a handful of 1-50 kb segments, inverted or tandem-duplicated
        n_inv = rng.poisson(inv_per_mbp * n_sites / site_bp)
        n_dup = rng.poisson(dup_per_mbp * n_sites / site_bp)
        events = []
        for kind, cnt in (("inv", n_inv), ("dup", n_dup)):
            for _e in range(cnt):
                seg = max(2, int(rng.integers(1_000, 50_000) / mean_len))
                if seg >= n_sites:
                    continue
                s = int(rng.integers(0, n_sites - seg))
                events.append((s, s + seg, kind, int(rng.integers(2, 6))))
        events.sort()
        idx_parts, cur = [], 0
        rev_mask_parts = []
        for s, e, kind, copies in events:
            if s < cur:
                continue  # overlapping event: skip
            idx_parts.append(np.arange(cur, s))
            rev_mask_parts.append(np.zeros(s - cur, dtype=np.uint8))
            if kind == "inv":
                idx_parts.append(np.arange(e - 1, s - 1, -1))
                rev_mask_parts.append(np.ones(e - s, dtype=np.uint8))
            else:
                idx_parts.append(np.tile(np.arange(s, e), copies))
                rev_mask_parts.append(np.zeros((e - s) * copies, dtype=np.uint8))
            cur = e
        idx_parts.append(np.arange(cur, n_sites))
        rev_mask_parts.append(np.zeros(n_sites - cur, dtype=np.uint8))
        idx = np.concatenate(idx_parts)
        step_node_parts.append(nodes[idx])
        step_rev_parts.append(rev[idx] | np.concatenate(rev_mask_parts))
        counts.append(idx.size)
    path_first = np.zeros(n_paths + 1, dtype=np.uint64)
    np.cumsum(np.array(counts, dtype=np.uint64), out=path_first[1:])
    step_node = np.concatenate(step_node_parts)
    step_rev = np.concatenate(step_rev_parts)
    g = FlatGraph(node_len, path_first, step_node, step_rev, None, [f"hap{i}" for i in range(n_paths)])
    if with_pos:
        lens = node_len[step_node].astype(np.uint64)
        csum = np.cumsum(lens, dtype=np.uint64)
        pos = np.empty_like(csum)
        pos[0] = 0
        pos[1:] = csum[:-1]
        base = pos[path_first[:-1].astype(np.int64)]
        g.step_pos = pos - np.repeat(base, np.array(counts))
    return g


# BASELINE config 5 (whole-human-pangenome scale: ~1e8 nodes, ~1e10 path steps, 100 haplotypes) and a small stand-in with the
# same structure for tests; generated RANK-LOCALLY (generate_sharded): no process ever holds the whole graph
SHARDED_PRESETS = {
    "c5": (83_000_000, 100),          # ~1.0e8 nodes, ~8.4e9 steps: 134 GB of step records -> 17 GB per GPU on 8
    "c5s": (2_000_000, 24),           # ~2.4e6 nodes, ~4.9e7 steps (tests, 2-GPU runs)
}


def _node_table(n_sites: int, seed: int, bubble_every: int):
    """the part of the graph every rank needs: alleles per site, node lengths, allele frequencies (same on every rank)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    k = np.ones(n_sites, dtype=np.int64)
    bub = np.arange(bubble_every // 2, n_sites, bubble_every)
    k[bub] = rng.integers(2, 5, size=bub.size)
    site_first = np.zeros(n_sites + 1, dtype=np.int64)
    np.cumsum(k, out=site_first[1:])
    N = int(site_first[-1])
    node_len = np.ones(N, dtype=np.uint32)
    longer = rng.random(N) >= 0.55
    node_len[longer] = 1 + np.minimum(rng.geometric(1.0 / 24.0, size=int(longer.sum())), 1023).astype(np.uint32)
    maf = rng.beta(0.5, 0.5, size=bub.size) * 0.5
    return k, bub, site_first, node_len, maf


def _walk(rng, n_sites, k, bub, site_first, maf, mean_len, inv_per_mbp, dup_per_mbp):
    """one haplotype: allele choice per bubble, then a few inverted / tandem-duplicated segments (as generate() does)"""
    site_bp = 1e6 / mean_len
    allele = np.zeros(n_sites, dtype=np.int64)
    u = rng.random(bub.size)
    minor = u < maf
    kk = k[bub] - 1
    allele[bub] = np.where(minor, 1 + np.minimum((u / np.maximum(maf, 1e-12) * kk).astype(np.int64), kk - 1), 0)
    nodes = (site_first[:-1] + allele).astype(np.uint32)
    del allele
    events = []
    for kind, cnt in (("inv", rng.poisson(inv_per_mbp * n_sites / site_bp)), ("dup", rng.poisson(dup_per_mbp * n_sites / site_bp))):
        for _e in range(cnt):
            seg = max(2, int(rng.integers(1_000, 50_000) / mean_len))
            if seg >= n_sites:
                continue
            s0 = int(rng.integers(0, n_sites - seg))
            events.append((s0, s0 + seg, kind, int(rng.integers(2, 6))))
    events.sort()
    node_parts, rev_parts, cur = [], [], 0
    for s0, e0, kind, copies in events:
        if s0 < cur:
            continue
        node_parts.append(nodes[cur:s0])
        rev_parts.append(np.zeros(s0 - cur, dtype=np.uint8))
        if kind == "inv":
            node_parts.append(nodes[s0:e0][::-1])
            rev_parts.append(np.ones(e0 - s0, dtype=np.uint8))
        else:
            node_parts.append(np.tile(nodes[s0:e0], copies))
            rev_parts.append(np.zeros((e0 - s0) * copies, dtype=np.uint8))
        cur = e0
    node_parts.append(nodes[cur:])
    rev_parts.append(np.zeros(n_sites - cur, dtype=np.uint8))
    return np.concatenate(node_parts), np.concatenate(rev_parts)


def generate_sharded(n_sites: int, n_paths: int, rank: int, n_ranks: int, seed: int = 42, bubble_every: int = 10,
                     inv_per_mbp: float = 0.5, dup_per_mbp: float = 0.2):
    """Rank-local generation of a path-sharded graph (BASELINE config 5): the node table is the same on every
    rank (derived from `seed`); path p is drawn from its OWN generator (seed, p), and rank r holds the paths p = r (mod
    n_ranks).  Returns (FlatGraph of this rank's paths with the whole node table, global path ids of those paths).
    The union over the ranks is one well-defined graph, whatever n_ranks is (tests/test_multirank_cpu.py)."""
    k, bub, site_first, node_len, maf = _node_table(n_sites, seed, bubble_every)
    mean_len = float(node_len.mean())
    mine = list(range(rank, n_paths, n_ranks))
    parts_n, parts_r, counts = [], [], []
    for p in mine:
        rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([seed, 0x5AD, p])))
        sn, sr = _walk(rng, n_sites, k, bub, site_first, maf, mean_len, inv_per_mbp, dup_per_mbp)
        parts_n.append(sn); parts_r.append(sr); counts.append(sn.size)
    path_first = np.zeros(len(mine) + 1, dtype=np.uint64)
    np.cumsum(np.array(counts, dtype=np.uint64), out=path_first[1:])
    step_node = np.concatenate(parts_n) if parts_n else np.zeros(0, dtype=np.uint32)
    step_rev = np.concatenate(parts_r) if parts_r else np.zeros(0, dtype=np.uint8)
    return FlatGraph(node_len, path_first, step_node, step_rev, None, [f"hap{p}" for p in mine]), mine


def preset(name: str, seed: int = 42, with_pos: bool = False) -> FlatGraph:
    n_sites, n_paths = PRESETS[name]
    return generate(n_sites, n_paths, seed=seed, with_pos=with_pos)


def write_gfa(g: FlatGraph, path: str) -> None:
    """GFA1 with placeholder sequences (only lengths matter to PG-SGD): lets the unmodified reference load a
    synthetic graph.  Edges are the adjacencies the paths use."""
    first = g.path_first_step.astype(np.int64)
    rev = g.step_rev if g.step_rev is not None else np.zeros(g.S, dtype=np.uint8)
    with open(path, "w") as f:
        f.write("H\tVN:Z:1.0\n")
        f.write("".join(f"S\t{i + 1}\t{'A' * int(ln)}\n" for i, ln in enumerate(g.node_len)))
        # unique (from handle, to handle) pairs over all consecutive steps
        h = (g.step_node.astype(np.uint64) << np.uint64(1)) | rev.astype(np.uint64)
        keep = np.ones(g.S, dtype=bool)
        keep[first[1:] - 1] = False  # the last step of a path has no successor
        keep = keep[:-1] if g.S else keep
        pairs = np.unique((h[:-1][keep[: g.S - 1]] << np.uint64(32)) | h[1:][keep[: g.S - 1]])
        u, v = pairs >> np.uint64(32), pairs & np.uint64(0xFFFFFFFF)
        sign = np.array(["+", "-"])
        f.write("".join(f"L\t{int(a >> 1) + 1}\t{sign[int(a & 1)]}\t{int(b >> 1) + 1}\t{sign[int(b & 1)]}\t0M\n" for a, b in zip(u, v)))
        for p in range(g.P):
            a, b = int(first[p]), int(first[p + 1])
            ids = (g.step_node[a:b].astype(np.int64) + 1).astype(str)
            steps = ",".join(np.char.add(ids, sign[rev[a:b]]))
            name = g.path_names[p] if p < len(g.path_names) else f"path{p}"
            f.write(f"P\t{name}\t{steps}\t*\n")
