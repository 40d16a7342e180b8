#!/bin/bash
# Round 2, call 33 (8 GPUs): the N=8 and N=4 bench lines (auto) on the final library (e2e with the caller-owned pinned result buffers).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 > gpurun_out/r02_c33_bench_c4_n8.json 2> gpurun_out/r02_c33_bench_c4_n8.err
timeout 300 $TR --nproc-per-node 4 --master-port 29572 bench.py --gpus 4 > gpurun_out/r02_c33_bench_c4_n4.json 2> gpurun_out/r02_c33_bench_c4_n4.err
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.loads(open(f"gpurun_out/r02_c33_bench_c4_n{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), {k: round(v, 4) for k, v in d["e2e"]["phases_rank0"].items()}, d["config"]["multi_mode"], d["quality"]["stress_final"])
    except Exception as ex:
        print(n, "failed", ex)
PY
tail -2 gpurun_out/r02_c33_bench_c4_n8.err
