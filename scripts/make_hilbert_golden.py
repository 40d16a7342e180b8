#!/usr/bin/env python
"""Golden values of the reference's Hilbert initialisation (`odgi layout -N h`, src/algorithms/hilbert.hpp:30-41 called with
n = 2 * node count, layout_main.cpp:287,312-318).  Authoring container only: compiles a 5-line probe that includes the
reference header where it lies and prints d2xy(n, d) for d < n."""
import hashlib
import json
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = r'''
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "algorithms/hilbert.hpp"
int main(int, char** argv) { uint64_t n = strtoull(argv[1], 0, 10); for (uint64_t d = 0; d < n; ++d) { uint64_t x, y; d2xy(n, d, &x, &y); printf("%lu %lu\n", x, y); } }
'''


def main():
    tmp = tempfile.mkdtemp()
    src, exe = os.path.join(tmp, "probe.cpp"), os.path.join(tmp, "probe")
    with open(src, "w") as f:
        f.write(PROBE)
    subprocess.run(["/usr/bin/g++", "-O1", "-I/root/reference/src", "-o", exe, src], check=True)
    gold = {}
    for n in (2, 8, 20, 30, 9910, 7502):   # 2 * nodes of note5-like, t, overlap, k, DRB1-3123, LPA
        out = subprocess.run([exe, str(n)], check=True, capture_output=True, text=True).stdout
        pts = [[int(v) for v in line.split()] for line in out.splitlines()]
        gold[str(n)] = {"sha256": hashlib.sha256(out.encode()).hexdigest(), "points": pts if n <= 30 else pts[:16]}
    with open(os.path.join(ROOT, "tests", "golden", "hilbert.json"), "w") as f:
        json.dump(gold, f)
    print("hilbert.json:", {k: v["sha256"][:12] for k, v in gold.items()})


if __name__ == "__main__":
    main()
