#!/bin/bash
# Round 2, call 19 (1 GPU): where does the e2e overhead go?  Engine-create phase times (PGSGD_TIMING) + bench line with pinned result buffers.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
PGSGD_TIMING=1 timeout 600 python bench.py --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_c19_bench.json 2> gpurun_out/r02_c19_bench.err
grep pgsgd_engine_create gpurun_out/r02_c19_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c19_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"])
PY
