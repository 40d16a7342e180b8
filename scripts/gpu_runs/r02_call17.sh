#!/bin/bash
# Round 2, call 17 (1 GPU): (a) the memory-traffic skeleton of a term (scripts/probes/gather_probe.cu): what does the access
# pattern alone allow?  (b) hub onset.  (c) one bench line on c4x (5e8 steps).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
  echo "== c4 shape, early iterations (76 % of partners outside the tile)"; timeout 200 scripts/probes/gather_probe 419 5.5 76
  echo "== c4 shape, cooling iterations (55 %)"; timeout 200 scripts/probes/gather_probe 419 5.5 55
  echo "== mid shape (coordinates fit L2), early"; timeout 200 scripts/probes/gather_probe 46 0.6 76
} > gpurun_out/r02_c17_gather_probe.log 2>&1
timeout 600 python scripts/gpu_exp_hub_onset.py > gpurun_out/r02_c17_hub_onset.log 2>&1
timeout 900 python bench.py --workload c4x --steps 10 --warmup 3 --no-reference-cuda > gpurun_out/r02_c17_bench_c4x.json 2> gpurun_out/r02_c17_bench_c4x.err
cat gpurun_out/r02_c17_gather_probe.log; tail -22 gpurun_out/r02_c17_hub_onset.log; cut -c1-600 gpurun_out/r02_c17_bench_c4x.json; tail -3 gpurun_out/r02_c17_bench_c4x.err
