#!/usr/bin/env python
"""Pin the oracle against the reference on RANDOM small graphs (authoring container only; nothing is committed but this
script): random node lengths, random walks with reverse-strand steps, 1-step paths, revisited nodes.  For each graph the
unmodified reference runs single-threaded with its -Deval_path_sgd trace hook (2D with the cooling switch, 1D), and the
oracle must replay the trace and the final fp64 coordinates bit for bit — the same procedure as scripts/pin_oracle.py."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import odgi_b200  # noqa: E402
import pin_oracle as po  # noqa: E402
from odgi_b200 import synth  # noqa: E402


def main():
    n_graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rng = np.random.default_rng(2024)
    with tempfile.TemporaryDirectory() as tmp:
        po.TEST, po.GOLD = tmp, os.path.join(tmp, "gold")     # fixtures of this run are thrown away
        os.makedirs(po.GOLD)
        for k in range(n_graphs):
            N, P = int(rng.integers(2, 80)), int(rng.integers(1, 7))
            node_len = rng.integers(1, 60, size=N).astype(np.uint32)
            counts = rng.integers(2, 120, size=P)
            if P > 1 and k % 3 == 0:
                counts[rng.integers(0, P)] = 1                # a 1-step path
            first = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
            S = int(first[-1])
            step_node = rng.integers(0, N, size=S).astype(np.uint32)
            step_rev = (rng.random(S) < 0.3).astype(np.uint8)
            name = f"fuzz{k}"
            synth.write_gfa(odgi_b200.FlatGraph(node_len, first, step_node, step_rev), os.path.join(tmp, name + ".gfa"))
            po.GRAPHS[name] = name + ".gfa"
            arrs = po.dump_graph(name, tmp)
            po.pin_2d(name, arrs, tmp, cooling_start=0.5, updates=2500, tag="cool")
            po.pin_1d(name, arrs, tmp, updates=2500)
            if k % 4 == 1:
                po.pin_1d(name, arrs, tmp, updates=2500, freeze_mod=2)    # `sort -H`: every second node stays put
    print(f"[fuzz] {n_graphs} random graphs: 2D and 1D traces and final coordinates bit-exact vs the reference")


if __name__ == "__main__":
    main()
