#!/bin/bash
# Round 2, call 29 (1 GPU): e2e again with the persistent staging buffers (3 bench runs: how much is box noise?)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python bench.py --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_c29_bench_$i.json 2>/dev/null; done
python - <<'PY'
import json
for i in (1, 2, 3):
    d = json.loads(open(f"gpurun_out/r02_c29_bench_{i}.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 4) for k, v in d["e2e"]["phases_rank0"].items()}, d["clocks"])
PY
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py -q -m gpu 2>&1 | tail -3
uptime
