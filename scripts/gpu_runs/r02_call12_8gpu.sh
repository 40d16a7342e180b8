#!/bin/bash
# Round 2, call 12 (8 GPUs): c4 at N=8 again (zeta tables cached: the host gap of the timed region), the shallow graph in the mode
# AUTO now picks for it (peer from the first iteration), c4 at N=4.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 > gpurun_out/r02_c12_bench_c4_n8.json 2> gpurun_out/r02_c12_bench_c4_n8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29562 scripts/gpu_multi_suite.py longthin --modes=peer,auto > gpurun_out/r02_c12_suite_longthin_n8.jsonl 2> gpurun_out/r02_c12_suite_longthin_n8.err
timeout 400 $TR --nproc-per-node 4 --master-port 29563 bench.py --gpus 4 --no-e2e > gpurun_out/r02_c12_bench_c4_n4.json 2> gpurun_out/r02_c12_bench_c4_n4.err
head -c 1800 gpurun_out/r02_c12_bench_c4_n8.json; echo; cat gpurun_out/r02_c12_suite_longthin_n8.jsonl; head -c 1200 gpurun_out/r02_c12_bench_c4_n4.json; echo; tail -3 gpurun_out/r02_c12_suite_longthin_n8.err
