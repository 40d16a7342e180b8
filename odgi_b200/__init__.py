"""odgi_b200 — B200-native (sm_100a) path-guided SGD for odgi layout / odgi sort.

The product is the C-ABI shared library built from odgi_b200/csrc (include/pgsgd.h); this package is
the thin Python binding used by tests and bench.py, plus graph I/O helpers.  The C++ host shim that
keeps odgi's own function signatures lives in odgi_b200/host/.
"""
from .capi import (Config, Engine, FlatGraph, PgsgdError, comm_unique_id, device_count, layout_2d, layout_2d_multi, layout_defaults,  # noqa: F401
                   schedule, sort_1d, sort_defaults, zetas)
from .graphio import assign_paths, graph_from_arrays, layout_init, load_graph_arrays, shard_paths  # noqa: F401
