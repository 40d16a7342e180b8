#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/refgpu_c4_prep.log 2>&1
import time, sys
sys.path.insert(0, '.')
from odgi_b200 import synth
t = time.time(); g = synth.preset("c4"); print("gen", time.time() - t, g.N, g.S, flush=True)
t = time.time(); synth.write_gfa(g, "/tmp/c4.gfa"); print("gfa", time.time() - t, flush=True)
PY
cat gpurun_out/refgpu_c4_prep.log
(cd /tmp && timeout 1800 /root/repo/oracle/_ref/ref_gpu_driver /tmp/c4.gfa - 10 64) > gpurun_out/refgpu_c4.json 2> gpurun_out/refgpu_c4.err
cat gpurun_out/refgpu_c4.json; tail -2 gpurun_out/refgpu_c4.err
