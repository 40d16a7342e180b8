"""Small invocations of every round-2 device path for compute-sanitizer (memcheck / racecheck):
pipelined tile kernel 2D + 1D (LDG+STS and TMA staging), stream kernel, stress / local stress, order with components, goodness,
.lay encoding with component stacking, GFA text ingest.  Sizes chosen so that a sanitizer run stays within a minute or two."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, synth  # noqa: E402

g = synth.generate(6_000, 5, seed=3, inv_per_mbp=40.0, dup_per_mbp=20.0)
X0, Y0 = odgi_b200.layout_init(g, 1)
with odgi_b200.Engine(g) as e:
    for flags in (0, capi.FLAG_TMA_STAGING | capi.FLAG_HALF_TILE, capi.FLAG_HALF_TILE, capi.FLAG_LEGACY_TILE):
        e.set_coords_2d(X0, Y0)
        st = e.run_range(capi.layout_defaults(g, iter_max=3, sampling=capi.SAMPLING_TILE, flags=flags), 2, 0, 3)
        assert st["term_updates"] == 3 * 10 * g.S, st
    e.set_coords_2d(X0, Y0)
    e.run_range(capi.layout_defaults(g, iter_max=2, sampling=capi.SAMPLING_STREAM), 2, 0, 2)
    print("2D ok", e.path_stress(2, 50_000, 1), e.local_stress(2, 50_000, 1))
    comp = (np.arange(g.N) % 3).astype(np.uint32)
    print("lay bytes", len(e.encode_lay(comp)), len(e.encode_lay()))
    for sampling, flags in ((capi.SAMPLING_TILE, 0), (capi.SAMPLING_TILE, capi.FLAG_TMA_STAGING | capi.FLAG_HALF_TILE), (capi.SAMPLING_STREAM, 0)):
        e.set_coords_1d(None)
        e.run_range(capi.sort_defaults(g, iter_max=3, sampling=sampling, flags=flags), 1, 0, 4)
    order = e.order_1d(comp)
    print("1D ok", e.path_stress(1, 50_000, 1), e.local_stress(1, 50_000, 1), e.sort_goodness(order)["mean_links_length_node"])
with tempfile.TemporaryDirectory() as tmp:
    p = os.path.join(tmp, "g.gfa")
    synth.write_gfa(g, p)
    with odgi_b200.Engine.from_gfa(p) as e2:
        assert e2.graph_stats()["step_count"] == g.S
        e2.set_coords_2d(X0, Y0)
        e2.run_range(capi.layout_defaults(g, iter_max=2, sampling=capi.SAMPLING_TILE), 2, 0, 2)
        print("ingest ok", e2.path_stress(2, 50_000, 1))
print("SANITIZE_SCRIPT_DONE")
