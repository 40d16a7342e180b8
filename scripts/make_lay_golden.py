#!/usr/bin/env python
"""Golden `.lay` files written by the reference's own algorithms::layout::Layout (through oracle/_ref/ref_driver lay_write).

Runs only in the authoring container (needs /root/reference and `make -C oracle/ref_build`).  For every coordinate set it
stores the input arrays and the bytes the reference wrote under tests/golden/, checks that odgi_b200/host/pgsgd writes the
same bytes and that both readers agree, and finally cross-checks the readers on the layout the reference ships with its
tests (test/DRB1-3123_unsorted.og.lay; not copied)."""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odgi_b200.arrays import read_arrays, write_arrays  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
CLI = os.path.join(ROOT, "odgi_b200", "host", "pgsgd")
GOLD = os.path.join(ROOT, "tests", "golden")


def cases():
    rng = np.random.default_rng(7)
    x = np.repeat(rng.normal(0, 100, 60), 3)
    return {
        "gauss300": (rng.normal(0, 1e4, 300), rng.normal(0, 50, 300)),
        "one": (np.array([3.5]), np.array([-2.25])),
        "n64": (rng.uniform(-1e6, 1e6, 64), rng.uniform(-1, 1, 64)),   # exactly one block of 128 values
        "n65": (rng.uniform(-1e6, 1e6, 65), rng.uniform(-1, 1, 65)),   # the second block holds two values
        "repeats": (x, x.copy()),                                       # equal neighbours: delta 0, coded as 2^64
        "monotone": (np.arange(400, dtype=np.float64) * 7.0, np.arange(400, dtype=np.float64) * 0.5 + 1000),
    }


def fuzz(n_cases: int):
    """random coordinate sets of six kinds (wide dynamic range, many equal values, sorted, subnormals, all equal, integers):
    writer bytes == reference bytes and reader == (X - min) + min, for every case"""
    tmp = tempfile.mkdtemp()
    rng = np.random.default_rng(123)
    for case in range(n_cases):
        n, kind = int(rng.integers(1, 700)), case % 6
        if kind == 0:
            X, Y = rng.normal(0, 10 ** rng.uniform(-5, 12), n), rng.normal(0, 10 ** rng.uniform(-5, 12), n)
        elif kind == 1:
            X, Y = np.round(rng.normal(0, 100, n)), np.round(rng.normal(0, 3, n))
        elif kind == 2:
            X, Y = np.sort(rng.uniform(0, 1e9, n)), np.zeros(n)
        elif kind == 3:
            X, Y = rng.uniform(-1, 1, n) * 1e-310, rng.uniform(-1, 1, n) * 1e-308
        elif kind == 4:
            X, Y = np.full(n, 7.25), np.full(n, 7.25)
        else:
            X, Y = rng.integers(-2 ** 40, 2 ** 40, n).astype(np.float64), rng.integers(0, 3, n).astype(np.float64)
        a, r, m, back = (os.path.join(tmp, f) for f in ("a.arr", "r.lay", "m.lay", "b.arr"))
        write_arrays(a, {"X": X.astype(np.float64), "Y": Y.astype(np.float64)})
        subprocess.run([REF, "lay_write", a, r], check=True)
        subprocess.run([CLI, "lay", "-c", a, "-o", m], check=True)
        assert open(r, "rb").read() == open(m, "rb").read(), (case, kind, n)
        subprocess.run([CLI, "lay", "-i", r, "-a", back], check=True)
        b, mv = read_arrays(back), min(X.min(), Y.min())
        assert np.array_equal(b["X"], (X - mv) + mv) and np.array_equal(b["Y"], (Y - mv) + mv), (case, kind, n)
    shutil.rmtree(tmp)
    print(f"[lay] fuzz: {n_cases} random coordinate sets, writer byte-identical to the reference and reader exact in every case")


def main():
    if "--fuzz" in sys.argv[1:]:
        return fuzz(int(sys.argv[sys.argv.index("--fuzz") + 1]))
    tmp = tempfile.mkdtemp()
    for name, (X, Y) in cases().items():
        arr = os.path.join(GOLD, f"lay_{name}.arr.gz")
        write_arrays(arr, {"X": X.astype(np.float64), "Y": Y.astype(np.float64)})
        plain = os.path.join(tmp, name + ".arr")
        write_arrays(plain, {"X": X.astype(np.float64), "Y": Y.astype(np.float64)})
        ref_lay = os.path.join(GOLD, f"lay_{name}.lay")
        subprocess.run([REF, "lay_write", plain, ref_lay], check=True)
        mine = os.path.join(tmp, name + ".lay")
        subprocess.run([CLI, "lay", "-c", plain, "-o", mine], check=True)
        assert open(mine, "rb").read() == open(ref_lay, "rb").read(), name
        back, refback = os.path.join(tmp, name + ".back.arr"), os.path.join(tmp, name + ".refback.arr")
        subprocess.run([CLI, "lay", "-i", ref_lay, "-a", back], check=True)
        subprocess.run([REF, "lay_read", ref_lay, refback], check=True)
        a, b = read_arrays(back), read_arrays(refback)
        assert np.array_equal(a["X"], b["X"]) and np.array_equal(a["Y"], b["Y"]), name
        print(f"[lay] {name}: {os.path.getsize(ref_lay)} bytes, writer byte-identical, readers agree")
    shipped = "/root/reference/test/DRB1-3123_unsorted.og.lay"
    back, refback = os.path.join(tmp, "shipped.arr"), os.path.join(tmp, "shipped.ref.arr")
    subprocess.run([CLI, "lay", "-i", shipped, "-a", back], check=True)
    subprocess.run([REF, "lay_read", shipped, refback], check=True)
    a, b = read_arrays(back), read_arrays(refback)
    assert np.array_equal(a["X"], b["X"]) and np.array_equal(a["Y"], b["Y"])
    print(f"[lay] reference test layout ({a['X'].size} points): readers agree")
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
