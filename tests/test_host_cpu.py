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


def test_gfa_reader_details(tmp_path):
    """S lines with '*' sequences and LN:i: tags, L lines and other records skipped, reverse steps, overlaps column absent."""
    gfa = tmp_path / "g.gfa"
    gfa.write_text("H\tVN:Z:1.0\n"
                   "S\t1\tACGT\n"
                   "S\t2\t*\tLN:i:7\n"
                   "S\t3\tGG\tRC:i:4\n"
                   "L\t1\t+\t2\t+\t0M\n"
                   "W\tsample\t1\tchr\t0\t9\t>1>2\n"
                   "P\tfwd\t1+,2+,3+\t*\n"
                   "P\trev\t3-,2-,1-\n")
    out = tmp_path / "g.arr"
    subprocess.run([CLI, "flatten", "-i", str(gfa), "-o", str(out)], check=True, capture_output=True)
    a = read_arrays(str(out))
    assert a["node_len"].tolist() == [4, 7, 2]
    assert a["path_first_step"].tolist() == [0, 3, 6]
    assert a["step_node"].tolist() == [0, 1, 2, 2, 1, 0]
    assert a["step_rev"].tolist() == [0, 0, 0, 1, 1, 1]
    assert a["step_pos"].tolist() == [0, 4, 11, 0, 2, 9]
    assert bytes(a["path_names"]).decode() == "fwd\nrev\n"


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` times the reference's CPU implementation (oracle/_ref when built, else the oracle port)
    and prints ONE JSON line with the contract's keys; it needs no GPU."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "M updates/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


LAY_CASES = ["gauss300", "one", "n64", "n65", "repeats", "monotone"]


@pytest.mark.parametrize("name", LAY_CASES)
def test_lay_writer_is_byte_identical_to_the_reference(golden_dir, tmp_path, name):
    """`.lay` (layout.cpp:43-61 = min_value + sdsl::enc_vector<> of bit-cast doubles) restated without sdsl:
    tests/golden/lay_*.lay were written by the reference's own Layout::serialize (scripts/make_lay_golden.py)."""
    from odgi_b200.arrays import write_arrays
    a = read_arrays(os.path.join(golden_dir, f"lay_{name}.arr.gz"))
    plain = tmp_path / "xy.arr"
    write_arrays(str(plain), {"X": a["X"], "Y": a["Y"]})
    out = tmp_path / "mine.lay"
    subprocess.run([CLI, "lay", "-c", str(plain), "-o", str(out)], check=True)
    with open(os.path.join(golden_dir, f"lay_{name}.lay"), "rb") as f:
        assert out.read_bytes() == f.read()


@pytest.mark.parametrize("name", LAY_CASES)
def test_lay_reader_returns_what_layout_get_x_returns(golden_dir, tmp_path, name):
    """Layout::get_x / get_y (layout.cpp:86-97): (X - min_value) + min_value, in that order of fp64 operations."""
    a = read_arrays(os.path.join(golden_dir, f"lay_{name}.arr.gz"))
    back, tsv = tmp_path / "back.arr", tmp_path / "back.tsv"
    subprocess.run([CLI, "lay", "-i", os.path.join(golden_dir, f"lay_{name}.lay"), "-a", str(back), "-T", str(tsv)], check=True)
    b = read_arrays(str(back))
    m = min(a["X"].min(), a["Y"].min())
    assert np.array_equal(b["X"], (a["X"] - m) + m) and np.array_equal(b["Y"], (a["Y"] - m) + m)
    rows = np.loadtxt(str(tsv), skiprows=1, ndmin=2)   # Layout::to_tsv: idx X Y with 16 significant digits
    assert rows.shape == (a["X"].size, 3) and np.array_equal(rows[:, 0], np.arange(a["X"].size))
    assert np.allclose(rows[:, 1], b["X"], rtol=1e-15, atol=0) and np.allclose(rows[:, 2], b["Y"], rtol=1e-15, atol=0)


def test_lay_reader_rejects_garbage(tmp_path):
    bad = tmp_path / "bad.lay"
    bad.write_bytes(b"\x00" * 7)
    r = subprocess.run([CLI, "lay", "-i", str(bad), "-a", str(tmp_path / "x.arr")], capture_output=True, text=True)
    assert r.returncode == 1 and "truncated" in r.stderr
    with open(os.path.join(ROOT, "tests", "golden", "lay_gauss300.lay"), "rb") as f:
        data = f.read()
    bad.write_bytes(data[: len(data) // 2])
    r = subprocess.run([CLI, "lay", "-i", str(bad), "-a", str(tmp_path / "x.arr")], capture_output=True, text=True)
    assert r.returncode == 1 and "lay:" in r.stderr


@pytest.mark.parametrize("name", ["t", "overlap", "DRB1-3123", "LPA"])
def test_hilbert_initialisation_matches_the_reference(golden_dir, golden_graphs, tmp_path, name):
    """`-N h`: d2xy(2N, 2r), d2xy(2N, 2r+1) per node rank r, with the reference's non-power-of-two side length
    (tests/golden/hilbert.json holds the reference function's output, scripts/make_hilbert_golden.py)."""
    import hashlib
    import json
    g = odgi_b200.graph_from_arrays(golden_graphs[name])
    gfa, arr = tmp_path / "g.gfa", tmp_path / "init.arr"
    synth.write_gfa(g, str(gfa))
    subprocess.run([CLI, "init", "-i", str(gfa), "-N", "h", "-a", str(arr)], check=True)
    b = read_arrays(str(arr))
    with open(os.path.join(golden_dir, "hilbert.json")) as f:
        gold = json.load(f)[str(2 * g.N)]
    text = "".join(f"{int(x)} {int(y)}\n" for x, y in zip(b["X"], b["Y"]))
    assert hashlib.sha256(text.encode()).hexdigest() == gold["sha256"]
    pts = np.array(gold["points"], dtype=np.float64)
    assert np.array_equal(b["X"][: len(pts)], pts[:, 0]) and np.array_equal(b["Y"][: len(pts)], pts[:, 1])


def test_default_initialisation_shape(golden_graphs, tmp_path):
    """`-N d` (layout_main.cpp:322-328): X = cumulative bp at both node ends, Y ~ N(0, sqrt(2N)); seeded runs repeat."""
    a = golden_graphs["DRB1-3123"]
    g = odgi_b200.graph_from_arrays(a)
    gfa = tmp_path / "g.gfa"
    synth.write_gfa(g, str(gfa))
    outs = []
    for k in range(2):
        arr = tmp_path / f"init{k}.arr"
        subprocess.run([CLI, "init", "-i", str(gfa), "-N", "d", "--init-seed", "42", "-a", str(arr)], check=True)
        outs.append(read_arrays(str(arr)))
    cs = np.cumsum(a["node_len"].astype(np.float64))
    assert np.array_equal(outs[0]["X"][1::2], cs) and np.array_equal(outs[0]["X"][2::2], cs[:-1]) and outs[0]["X"][0] == 0
    assert np.array_equal(outs[0]["Y"], outs[1]["Y"]) and abs(outs[0]["Y"].std() / np.sqrt(2 * g.N) - 1) < 0.05
    r = subprocess.run([CLI, "init", "-i", str(gfa), "-N", "z", "-a", str(tmp_path / "z.arr")], capture_output=True, text=True)
    assert r.returncode == 1 and "unknown layout initialization" in r.stderr


REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


@pytest.mark.skipif(not os.path.exists(REF_DRIVER), reason="oracle/_ref/ref_driver not built (needs the reference tree at build time)")
@pytest.mark.parametrize("name", ["note5", "overlap", "DRB1-3123", "LPA"])
def test_shim_flatten_template_on_the_reference_graph_type(golden_graphs, tmp_path, name):
    """pgsgd::flatten_handle_graph instantiated on the reference's own graph_t (exactly what odgi_shim.cpp does), run inside
    the reference driver: equals the driver's walk, the XP accessors the CPU workers use, and the committed fixture."""
    a = golden_graphs[name]
    gfa, out = tmp_path / "g.gfa", tmp_path / "dump.arr"
    synth.write_gfa(odgi_b200.graph_from_arrays(a), str(gfa))
    subprocess.run([REF_DRIVER, "dump", str(gfa), str(out)], check=True, capture_output=True, cwd=str(tmp_path))
    d = read_arrays(str(out))
    for k in ("node_len", "path_first_step", "step_node", "step_rev", "step_pos"):
        assert np.array_equal(d["shim_" + k], d[k]) and np.array_equal(d["shim_" + k], a[k]), k
    assert np.array_equal(d["shim_step_pos"], d["xp_position_of_step"])
    assert np.array_equal((d["shim_step_node"].astype(np.uint64) << np.uint64(1)) | d["shim_step_rev"], d["xp_handle_of_step"])


def test_bench_line_assembles_with_a_stand_in_engine(monkeypatch, capfd):
    """bench.py's own arm needs a GPU; this runs its whole control flow on the CPU with the engine replaced by a stand-in that
    returns plausible statistics, so that the JSON contract (keys, units, one line on stdout) is guarded without a device."""
    import importlib
    import json
    import sys
    import types
    import torch
    from odgi_b200 import capi
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")

    class StandIn:
        device_bytes = 123

        def __init__(self, g, device=0):
            self.g, self.it = g, 0

        def set_coords_2d(self, X, Y):
            self.X, self.Y = np.array(X, dtype=np.float64), np.array(Y, dtype=np.float64)

        def path_stress(self, dims, pairs, seed):
            return 1.0 / (1 + self.it)

        def local_stress(self, dims, pairs, seed):
            return 2.0 / (1 + self.it)

        def run_range(self, cfg, dims, lo, hi):
            n = max(0, min(hi, cfg.iter_max) - lo)
            self.it += n
            return {"iterations_run": n, "kernel_launches": n, "term_updates": n * cfg.min_term_updates, "seconds_iterations": 1e-3 * n,
                    "h2d_bytes": 1000, "seconds_upload": 0.0}

        def get_coords_2d(self, out=None):
            if out is not None:
                out[0][...] = self.X
                out[1][...] = self.Y
                return out
            return self.X, self.Y

        def close(self):
            pass

    monkeypatch.setattr(bench, "pinned_like", lambda a: (np.array(a), None))
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda i: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(odgi_b200, "device_count", lambda: 1)
    monkeypatch.setattr(odgi_b200, "Engine", StandIn)
    # the DEFAULT workload name (its recorded-context files under profiles/ are read at the end), on a tiny graph
    real_workload = bench.make_workload
    monkeypatch.setattr(bench, "make_workload", lambda name: real_workload("tiny"))
    # (the live reference-CUDA-kernel leg launches oracle/_ref/ref_gpu_driver on a GPU: switched off here)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "4", "--warmup", "3", "--no-cpu-baseline", "--no-reference-cuda"])
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    bench.main()
    out = capfd.readouterr().out.strip().splitlines()
    assert len(out) == 1, out
    line = json.loads(out[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "gpu_launches", "clocks", "e2e", "roofline"):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["steps"] == 4 and line["warmup"] == 3 and line["gpu_launches"] == 4 and line["vs_baseline"] is None
    assert line["config"]["workload"] == "c4" and "model" not in line["config"]
    assert "reference_cuda_kernel" not in line and line["roofline"]["traffic"]
    assert line["config"]["timed"].startswith("CUDA events")
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "phases_rank0"}
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and line["roofline"]["bound"] == "hbm"
    assert abs(line["value"] - 4 * line["config"]["updates_per_step"] / 4e-3 / 1e6) < 1e-6 * line["value"]


def test_sort_target_paths_and_use_paths_preparation(golden_graphs, tmp_path):
    """`pgsgd sort -H targets -f use`: what reaches the GPU (--prepared-out, no device needed) against an independent restatement
    of sort_graph_by_target_paths (sort_main.cpp:266-311: nodes of the target paths first, in first-visit order; the rest in
    id order; the first ref_nodes ranks frozen) and of the -f parameter derivation (sort_main.cpp:355-412)."""
    a = golden_graphs["DRB1-3123"]
    g = odgi_b200.graph_from_arrays(a)
    names = bytes(a["path_names"]).decode().split("\n")[:-1]
    gfa, prep = tmp_path / "g.gfa", tmp_path / "prep.arr"
    synth.write_gfa(g, str(gfa))
    targets, use = [names[5], names[2], "not-in-the-graph"], [names[1], names[7], names[3]]
    (tmp_path / "targets.txt").write_text("\n".join(targets) + "\n")
    (tmp_path / "use.txt").write_text("\n".join(use) + "\n")
    r = subprocess.run([CLI, "sort", "-i", str(gfa), "-o", str(tmp_path / "order.txt"), "-Y", "-H", str(tmp_path / "targets.txt"),
                        "-f", str(tmp_path / "use.txt"), "--prepared-out", str(prep)], capture_output=True, text=True)
    assert r.returncode == 0 and "found 2/3 paths to consider" in r.stderr, r.stderr
    p = read_arrays(str(prep))
    first, node = a["path_first_step"].astype(np.int64), a["step_node"].astype(np.int64)
    order, seen = [], np.zeros(g.N, dtype=bool)
    for name in targets[:2]:
        k = names.index(name)
        for n in node[first[k]:first[k + 1]]:
            if not seen[n]:
                seen[n] = True
                order.append(n)
    ref_nodes = len(order)
    order += [n for n in range(g.N) if not seen[n]]
    order = np.array(order)
    new_of_old = np.empty(g.N, dtype=np.int64)
    new_of_old[order] = np.arange(g.N)
    assert np.array_equal(p["old_of_new"], order)
    assert np.array_equal(p["frozen"], (np.arange(g.N) < ref_nodes).astype(np.uint8)) and 0 < ref_nodes < g.N
    assert np.array_equal(p["node_len"], a["node_len"][order]) and np.array_equal(p["step_node"], new_of_old[node])
    assert np.array_equal(p["step_pos"], a["step_pos"]) and np.array_equal(p["step_rev"], a["step_rev"])
    ks = [names.index(n) for n in use]
    steps = [int(first[k + 1] - first[k]) for k in ks]
    bps = [int(a["step_pos"][first[k + 1] - 1]) + int(a["node_len"][node[first[k + 1] - 1]]) for k in ks]
    assert p["min_term_updates"][0] == sum(steps) and p["space"][0] == max(bps) and p["eta_max"][0] == float(max(steps)) ** 2
    assert p["space_max"][0] == 100 and p["space_quantization_step"][0] == max(2, -(-(max(bps) - 100) // 1))
    # duplicates in the target list are an error, as in the reference
    (tmp_path / "dup.txt").write_text(names[0] + "\n" + names[0] + "\n")
    r = subprocess.run([CLI, "sort", "-i", str(gfa), "-o", str(tmp_path / "o.txt"), "-Y", "-H", str(tmp_path / "dup.txt"), "--prepared-out", str(prep)],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "duplicated path names" in r.stderr


def test_gfa_line_scan_is_the_same_for_any_number_of_pieces(golden_graphs, tmp_path):
    """the device ingest's host half (scan_gfa: line boundaries, S lengths, P fields) cuts the file into pieces for host threads:
    what it finds must not depend on the cuts (PGSGD_SCAN_THREADS forces the piece count, also on a small file)"""
    import json
    g = odgi_b200.graph_from_arrays(golden_graphs["DRB1-3123"])
    gfa = tmp_path / "g.gfa"
    synth.write_gfa(g, str(gfa))
    outs = []
    for t in ("1", "2", "5", "64"):
        r = subprocess.run([CLI, "scan", "-i", str(gfa)], capture_output=True, text=True, env={**os.environ, "PGSGD_SCAN_THREADS": t})
        assert r.returncode == 0, r.stderr
        d = json.loads(r.stdout)
        d.pop("seconds")
        outs.append(d)
    assert all(o == outs[0] for o in outs[1:]), outs
    assert outs[0]["nodes"] == g.N and outs[0]["paths"] == g.P and outs[0]["bp"] == int(g.node_len.sum())
    bad = tmp_path / "bad.gfa"
    bad.write_text("H\tVN:Z:1.0\nS\t1\tACGT\nS\t3\tA\nP\tp\t1+,3+\t*\n")
    r = subprocess.run([CLI, "scan", "-i", str(bad)], capture_output=True, text=True, env={**os.environ, "PGSGD_SCAN_THREADS": "3"})
    assert r.returncode == 1 and "not optimized" in r.stderr
