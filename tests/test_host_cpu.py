"""Host-side C++ (odgi_b200/host): the standalone GFA flattening against the reference's own tables."""
import os
import subprocess

import numpy as np
import pytest

from odgi_b200 import synth
from odgi_b200.arrays import read_arrays
import odgi_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "odgi_b200", "host", "pgsgd")


@pytest.fixture(scope="module", autouse=True)
def _built():
    from odgi_b200.build import build_native
    build_native()
    subprocess.run(["make", "-C", os.path.join(ROOT, "odgi_b200", "host")], check=True, capture_output=True)


@pytest.mark.parametrize("name", ["note5", "t", "DRB1-3123", "chr6.C4"])
def test_gfa_flatten_matches_reference_tables(golden_graphs, tmp_path, name):
    """GFA -> our flattening == the arrays dumped from the reference (graph_t walk; step_pos == XP get_position_of_step)."""
    a = golden_graphs[name]
    g = odgi_b200.graph_from_arrays(a)
    gfa = tmp_path / f"{name}.gfa"
    synth.write_gfa(g, str(gfa))
    out = tmp_path / "flat.arr"
    subprocess.run([CLI, "flatten", "-i", str(gfa), "-o", str(out)], check=True, capture_output=True)
    b = read_arrays(str(out))
    for k in ("node_len", "path_first_step", "step_node", "step_rev", "step_pos"):
        assert np.array_equal(a[k], b[k]), k


def test_cli_refuses_without_gpu_flag_and_on_unoptimized_graph(tmp_path):
    gfa = tmp_path / "bad.gfa"
    gfa.write_text("H\tVN:Z:1.0\nS\t1\tACGT\nS\t3\tA\nP\tp\t1+,3+\t*\n")  # ids 1,3: not compacted
    r = subprocess.run([CLI, "layout", "-i", str(gfa), "-T", "-"], capture_output=True, text=True)
    assert r.returncode == 1 and "--gpu" in r.stderr
    r = subprocess.run([CLI, "flatten", "-i", str(gfa), "-o", str(tmp_path / "x.arr")], capture_output=True, text=True)
    assert r.returncode == 1 and "not optimized" in r.stderr
