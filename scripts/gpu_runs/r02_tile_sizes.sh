#!/bin/bash
# Round 2 experiment: tile size of the tile-sampling kernel (build-time knob), c4 bench line per size.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_runs/r02_tile_sizes.sh'
# The library is rebuilt ON the box (same image, nvcc present) and left at the default size at the end.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for T in 1024 4096 2048; do
    PGSGD_TILE_STEPS=$T python -c "import odgi_b200.build as b; b.build_native(force=True)" > /dev/null 2>&1
    timeout 400 python bench.py --no-e2e --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/r02_tile_${T}.json 2> gpurun_out/r02_tile_${T}.err
    python - <<PY
import json
d = json.load(open("gpurun_out/r02_tile_${T}.json"))
print("tile ${T}:", round(d["value"] / 1e3, 2), "G updates/s, stress", d.get("quality", {}).get("stress_final"))
PY
done
