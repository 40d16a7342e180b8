#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
# reference CUDA kernel (recompiled for sm_100a) vs ours on the same synthetic graph (200k sites x 90 paths)
python - <<'PY' > gpurun_out/refgpu_prep.log 2>&1
import time, sys
sys.path.insert(0, '.')
from odgi_b200 import synth
t = time.time(); g = synth.generate(200_000, 90, seed=42); print("gen", time.time() - t, g.N, g.S)
t = time.time(); synth.write_gfa(g, "/tmp/refmid.gfa"); print("gfa", time.time() - t)
PY
(cd /tmp && timeout 1500 /root/repo/oracle/_ref/ref_gpu_driver /tmp/refmid.gfa - 30 32) > gpurun_out/refgpu_refmid.json 2> gpurun_out/refgpu_refmid.err
cat gpurun_out/refgpu_refmid.json
python - <<'PY' > gpurun_out/ours_refmid.log 2>&1
import sys
sys.path.insert(0, '.')
import odgi_b200
from odgi_b200 import capi, synth
g = synth.generate(200_000, 90, seed=42)
X0, Y0 = odgi_b200.layout_init(g, 42)
with odgi_b200.Engine(g) as e:
    for flags in (0, 1):
        for batch in (1, 4):
            cd = capi.layout_defaults(g, batch=batch, flags=flags)
            e.set_coords_2d(X0, Y0)
            st = e.run_2d(cd)
            print(f"ours refmid flags={flags} batch={batch}: {st['term_updates']/st['seconds_iterations']/1e9:.2f} G updates/s over the full 30-iteration run ({st['seconds_iterations']:.3f} s)", flush=True)
PY
cat gpurun_out/ours_refmid.log
timeout 1200 python bench.py --workload c4 --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 600 python scripts/gpu_exp_occupancy.py c4 > gpurun_out/exp_occ_c4.log 2>&1; tail -40 gpurun_out/exp_occ_c4.log
cut -c1-200 gpurun_out/bench_c4.json
