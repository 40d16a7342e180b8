"""GPU experiment (round 2): the final far / local stress of the REFERENCE'S OWN CUDA PATH (src/cuda/layout.cu compiled unmodified
for sm_100a, oracle/_ref/ref_gpu_driver) on the scale graphs, from the same injected initial layout as the CPU reference bands
(tests/golden/stress_reference_scale.json) — the implementation this library drops in for.  Its worker seeds are fixed
(layout.cu:29), so its runs differ by GPU timing only.  Evaluated with the device stress readout (== the oracle's definition).

  python scripts/gpu_exp_refcuda_stress.py <mid|longthin> [repeats=8]"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200  # noqa: E402
from odgi_b200 import capi, graphio, synth  # noqa: E402
from odgi_b200.arrays import read_arrays, write_arrays  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mid"
repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 8
g = synth.preset(name) if name in synth.PRESETS else synth.generate(3_000_000, 6, seed=42)
X0, Y0 = odgi_b200.layout_init(g, 42)
tmp = os.environ.get("TMPDIR", "/tmp")
arr, gfa, init, out = (os.path.join(tmp, f"{name}.{x}") for x in ("arr", "gfa", "init.arr", "refcuda"))
graphio.save_graph_arrays(arr, g)
subprocess.run([os.path.join(ROOT, "scripts", "probes", "arr2gfa"), arr, gfa], check=True, capture_output=True)
write_arrays(init, {"X": X0, "Y": Y0})
r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_gpu_driver"), gfa, out, "30", "32", "10", init, str(repeats)], capture_output=True, text=True, cwd=tmp)
print(r.stdout[-1500:], r.stderr[-500:], flush=True)
far, loc = [], []
with odgi_b200.Engine(g) as e:
    for k in range(repeats):
        a = read_arrays(f"{out}.run{k}.arr")
        e.set_coords_2d(a["X"], a["Y"])
        far.append(e.path_stress(2, 4_000_000, 12345)); loc.append(e.local_stress(2, 4_000_000, 12345))
    print(json.dumps({"graph": name, "impl": "reference src/cuda/layout.cu", "far": far, "local": loc, "far_mean": float(np.mean(far)), "far_sd": float(np.std(far, ddof=1)),
                      "local_mean": float(np.mean(loc)), "local_sd": float(np.std(loc, ddof=1))}), flush=True)
    for sname, sampling in (("tile", capi.SAMPLING_TILE), ("stream", capi.SAMPLING_STREAM)):
        f2, l2 = [], []
        for seed in [9399220, 1234567, 42] + [1000003 * (k + 1) for k in range(max(0, repeats - 3))]:
            e.set_coords_2d(X0, Y0)
            e.run_2d(capi.layout_defaults(g, seed=seed, sampling=sampling))
            f2.append(e.path_stress(2, 4_000_000, 12345)); l2.append(e.local_stress(2, 4_000_000, 12345))
        print(json.dumps({"graph": name, "impl": f"this library, {sname} sampling", "far": f2, "local": l2, "far_mean": float(np.mean(f2)), "far_sd": float(np.std(f2, ddof=1)),
                          "local_mean": float(np.mean(l2)), "local_sd": float(np.std(l2, ddof=1))}), flush=True)
