"""GFA P-line step lists parsed on the device (pgsgd_engine_create_from_gfa_paths) == the flattened fixtures
(which equal the reference's own graph walk and XP tables, tests/test_oracle_pinned.py): same step records, hence identical
sampler output and — single stream, strict order — bit-identical coordinates."""
import os

import numpy as np
import pytest

import odgi_b200
from odgi_b200 import capi, synth
from odgi_b200.arrays import read_arrays

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["DRB1-3123", "chr6.C4", "note5", "overlap", "k"])
def test_engine_from_gfa_text_equals_engine_from_arrays(tmp_path, name):
    a = read_arrays(os.path.join(GOLDEN, f"{name}.graph.arr.gz"))
    g = odgi_b200.graph_from_arrays(a)
    gfa = tmp_path / "g.gfa"
    synth.write_gfa(g, str(gfa))
    kw = dict(iter_max=3, min_term_updates=3000, eta_max=1000.0, n_streams=1, batch=1, sampling=capi.SAMPLING_STREAM)
    X0, Y0 = odgi_b200.layout_init(g, seed=5)
    with odgi_b200.Engine(g) as e1, odgi_b200.Engine.from_gfa(str(gfa)) as e2:
        st = e2.graph_stats()
        assert (e2.g.N, e2.g.P, st["step_count"]) == (g.N, g.P, g.S)
        assert st["max_path_steps"] == g.max_path_steps and st["max_path_bp"] == g.max_path_bp
        cfg = capi.layout_defaults(g, **kw)
        t1 = e1.sample_terms(cfg, 2, False, 3000)
        t2 = e2.sample_terms(cfg, 2, False, 3000)
        for k in t1:
            assert np.array_equal(t1[k], t2[k]), k          # nodes, orientations, bp positions of the sampled steps
        for e in (e1, e2):
            e.set_coords_2d(X0, Y0)
            e.run_2d(cfg)
        assert np.array_equal(e1.get_coords_2d()[0], e2.get_coords_2d()[0])


def test_engine_from_gfa_text_on_a_graph_beyond_one_chunk(tmp_path):
    """many 4 KB text chunks, ids of every width, reverse steps, CRLF-free; the step arrays themselves are compared through the
    device-evaluated stress of a fixed layout (depends on every step's node, orientation and position)"""
    g = synth.generate(60_000, 5, seed=9, inv_per_mbp=30.0, dup_per_mbp=10.0)
    gfa = tmp_path / "big.gfa"
    synth.write_gfa(g, str(gfa))
    X0, Y0 = odgi_b200.layout_init(g, seed=1)
    with odgi_b200.Engine(g) as e1, odgi_b200.Engine.from_gfa(str(gfa)) as e2:
        assert e2.graph_stats()["step_count"] == g.S
        for e in (e1, e2):
            e.set_coords_2d(X0, Y0)
        assert e1.path_stress(2, 400_000, 3) == e2.path_stress(2, 400_000, 3)
        assert e1.local_stress(2, 400_000, 4) == e2.local_stress(2, 400_000, 4)


def test_staged_text_upload_with_many_chunks_and_threads(tmp_path, monkeypatch):
    """the step lists go host -> device through pinned staging buffers filled by host threads, 16 MB chunks of the packed range;
    PGSGD_UPLOAD_CHUNK / PGSGD_UPLOAD_THREADS force many chunks (boundaries inside fields and between them) on a small file"""
    g = synth.generate(20_000, 7, seed=4, inv_per_mbp=30.0, dup_per_mbp=10.0)
    gfa = tmp_path / "g.gfa"
    synth.write_gfa(g, str(gfa))
    X0, Y0 = odgi_b200.layout_init(g, seed=1)
    with odgi_b200.Engine(g) as e1:
        e1.set_coords_2d(X0, Y0)
        want = (e1.path_stress(2, 300_000, 3), e1.local_stress(2, 300_000, 4))
    for chunk, threads in (("4096", "3"), ("1000", "5"), ("65536", "1")):
        monkeypatch.setenv("PGSGD_UPLOAD_CHUNK", chunk)
        monkeypatch.setenv("PGSGD_UPLOAD_THREADS", threads)
        with odgi_b200.Engine.from_gfa(str(gfa)) as e2:
            assert e2.graph_stats()["step_count"] == g.S
            e2.set_coords_2d(X0, Y0)
            assert (e2.path_stress(2, 300_000, 3), e2.local_stress(2, 300_000, 4)) == want, (chunk, threads)


def test_engine_from_gfa_text_rejects_bad_ids(tmp_path):
    bad = tmp_path / "bad.gfa"
    bad.write_text("H\tVN:Z:1.0\nS\t1\tACGT\nS\t2\tA\nP\tx\t1+,3+\t*\n")   # node 3 does not exist
    with pytest.raises(capi.PgsgdError):
        odgi_b200.Engine.from_gfa(str(bad))
