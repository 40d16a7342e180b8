#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/refgpu_mid_prep.log 2>&1
import time, sys
sys.path.insert(0, '.')
from odgi_b200 import synth
t = time.time(); g = synth.preset("mid"); print("gen", time.time() - t, g.N, g.S)
t = time.time(); synth.write_gfa(g, "/tmp/mid.gfa"); print("gfa", time.time() - t)
PY
cat gpurun_out/refgpu_mid_prep.log
(cd /tmp && timeout 1500 /root/repo/oracle/_ref/ref_gpu_driver /tmp/mid.gfa - 30 64) > gpurun_out/refgpu_mid.json 2> gpurun_out/refgpu_mid.err
cat gpurun_out/refgpu_mid.json
timeout 600 python scripts/gpu_exp_tile.py mid > gpurun_out/exp_tile_mid2.log 2>&1; cat gpurun_out/exp_tile_mid2.log
