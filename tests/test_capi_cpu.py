"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/pgsgd.h declares, its
pure-host helpers agree bit-for-bit with the oracle, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import odgi_b200
from odgi_b200 import capi
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    from odgi_b200.build import build_native
    build_native()


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pgsgd.h")).read()
    declared = set(re.findall(r"\b(pgsgd_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    L = C.CDLL(capi.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s
    assert L.pgsgd_version() == capi.ABI_VERSION


def test_struct_sizes_match_header():
    assert C.sizeof(capi.ConfigC) == 120
    assert C.sizeof(capi.GraphView) == 64
    assert C.sizeof(capi.StatsC) == 104


@pytest.mark.parametrize("kw", [dict(iter_max=30, eta_max=3100.0 ** 2, eps=0.01), dict(iter_max=100, eta_max=21901.0 ** 2, eps=0.01),
                                dict(iter_max=10, eta_max=1000.0, eps=0.5, iter_with_max_learning_rate=3)])
def test_schedule_matches_oracle(kw):
    a = capi.schedule(capi.Config(**kw))
    b = orc.schedule(orc.Config(**kw))
    assert np.array_equal(a, b)


# (1000, 1000, ...) and (100, 100, 2, ...): space == space_max, where the reference's loop stores one element past its table
# (path_sgd_layout.cpp:92-95) — here the store is dropped; a longest path of exactly 1000 steps reaches it
@pytest.mark.parametrize("space,space_max,q,theta", [(3100, 1000, 100, 0.99), (337324, 100, 337224, 0.99), (50, 1000, 100, 0.99), (5000, 10, 7, 0.5),
                                                     (1000, 1000, 100, 0.99), (100, 100, 2, 0.99), (7, 7, 100, 0.5)])
def test_zetas_match_oracle(space, space_max, q, theta):
    kw = dict(space=space, space_max=space_max, space_quantization_step=q, theta=theta)
    a = capi.zetas(capi.Config(**kw))
    b = orc.zetas(orc.Config(**kw))
    assert np.array_equal(a, b)


def test_no_cpu_fallback(golden_graphs):
    """Without a CUDA device the compute entry points must fail, not silently compute on the host."""
    if odgi_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    g = odgi_b200.graph_from_arrays(golden_graphs["t"])
    X, Y = odgi_b200.layout_init(g)
    with pytest.raises(odgi_b200.PgsgdError) as ei:
        odgi_b200.layout_2d(g, odgi_b200.layout_defaults(g), X, Y)
    assert ei.value.code == -2 and "no usable CUDA device" in str(ei.value)


def test_argument_validation(golden_graphs):
    a = golden_graphs["t"]
    bad = odgi_b200.FlatGraph(a["node_len"], np.array([1, 10], dtype=np.uint64), a["step_node"])
    with pytest.raises(odgi_b200.PgsgdError) as ei:
        odgi_b200.Engine(bad)
    assert ei.value.code == -1


def test_abi_version_is_one_number_everywhere():
    """include/pgsgd.h, the ctypes binding and the loaded library agree (a stale caller would pass short structs)."""
    import re
    with open(os.path.join(ROOT, "include", "pgsgd.h")) as f:
        m = re.search(r"#define\s+PGSGD_VERSION\s+(\d+)", f.read())
    assert m and int(m.group(1)) == capi.ABI_VERSION == capi.lib().pgsgd_version()


def test_header_is_plain_c99_and_links(tmp_path):
    """include/pgsgd.h is the drop-in boundary for ANY FFI: it must compile as strict C99 and link against the library."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "pgsgd.h"\n#include <stdio.h>\n'
                   'int main(void) { pgsgd_config c; pgsgd_graph_view v; pgsgd_stats s;\n'
                   '  printf("%d %d %d\\n", (int) sizeof c, (int) sizeof v, (int) sizeof s);\n'
                   '  return pgsgd_version() == PGSGD_VERSION && pgsgd_last_error() ? 0 : 1; }\n')
    exe = tmp_path / "hdr"
    libdir = os.path.join(ROOT, "odgi_b200")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lpgsgd_b200", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split() == ["120", "64", "104"]
