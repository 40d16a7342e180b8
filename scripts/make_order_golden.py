#!/usr/bin/env python
"""Golden fixtures for path_linear_sgd_order (path_sgd.cpp:503-684): X and the node order of ONE reference
run (`ref_driver sort order=1`: the order from path_linear_sgd_order, X recovered from the reference's own 1D .lay output
of the same call), on
  * multi3: a 40-node graph with three weak components whose node ids interleave (so that their X ranges overlap and the
    component key of the sort is observable), written here;
  * DRB1-3123 (one component).
Stores tests/golden/order_<name>.json {gfa (multi3 only), X (hex), order (handle integers)}.
Authoring container only (needs oracle/_ref)."""
import json
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odgi_b200.arrays import read_arrays  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
GOLD = os.path.join(ROOT, "tests", "golden")


def multi3_gfa() -> str:
    rnd = random.Random(11)
    ids = list(range(1, 41))
    comps = [[i for i in ids if i % 2 == 1], [i for i in ids if i % 2 == 0 and i <= 24], [i for i in ids if i % 2 == 0 and i > 24]]
    lines = ["H\tVN:Z:1.0"]
    for i in ids:
        lines.append(f"S\t{i}\t" + "".join(rnd.choice("ACGT") for _ in range(rnd.choice([1, 1, 2, 3, 8, 15]))))
    edges, paths = set(), []
    for ci, c in enumerate(comps):
        for pi in range(3):
            walk = [v for v in c if rnd.random() < 0.8 or v in (c[0], c[-1])]
            steps = [(v, "+") for v in walk]
            if len(steps) > 4 and pi == 1:
                k = len(steps) // 2
                steps[k] = (steps[k][0], "-")
            paths.append((f"c{ci}p{pi}", steps))
            for a, b in zip(steps, steps[1:]):
                edges.add((a[0], a[1], b[0], b[1]))
    for name, steps in paths:
        lines.append(f"P\t{name}\t" + ",".join(f"{v}{o}" for v, o in steps) + "\t*")
    for a, ao, b, bo in sorted(edges):
        lines.append(f"L\t{a}\t{ao}\t{b}\t{bo}\t0M")
    return "\n".join(lines) + "\n"


def run(gfa_path: str):
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "o.arr")
        subprocess.run([REF, "sort", gfa_path, out, "threads=1", "order=1"], check=True, cwd=tmp, capture_output=True)
        r = read_arrays(out)
        return [float(v).hex() for v in r["X"]], [int(v) for v in r["order"]]


def main():
    with tempfile.TemporaryDirectory() as tmp:
        text = multi3_gfa()
        p = os.path.join(tmp, "multi3.gfa")
        with open(p, "w") as f:
            f.write(text)
        X, order = run(p)
        with open(os.path.join(GOLD, "order_multi3.json"), "w") as f:
            json.dump({"gfa": text, "X": X, "order": order}, f, indent=0)
    X, order = run("/root/reference/test/DRB1-3123.gfa")
    with open(os.path.join(GOLD, "order_DRB1-3123.json"), "w") as f:
        json.dump({"X": X, "order": order}, f)
    print("ok")


if __name__ == "__main__":
    main()


# The flattened DRB1-3123_unsorted graph (the input of the reference's sorting tutorial, docs/rst/tutorials/sort_layout.rst)
# was added with:  oracle/_ref/ref_driver dump /root/reference/test/DRB1-3123_unsorted.gfa u.arr  -> keep node_len,
# path_first_step, step_node, step_rev, step_pos -> gzip -> tests/golden/DRB1-3123_unsorted.graph.arr.gz
