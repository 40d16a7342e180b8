#!/bin/bash
# Round 2, call 7 (8 GPUs): the official bench line at N=8 (c4, auto), the mode suite on c4 / mid / longthin with stress
# readouts, and BASELINE config 5 (rank-locally generated, path-sharded).
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_runs/r02_call7_8gpu.sh'
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29551 bench.py --gpus 8 > gpurun_out/r02_c7_bench_c4_n8.json 2> gpurun_out/r02_c7_bench_c4_n8.err
timeout 500 $TR --master-port 29552 scripts/gpu_multi_suite.py c4 mid longthin > gpurun_out/r02_c7_suite_n8.jsonl 2> gpurun_out/r02_c7_suite_n8.err
free -g > gpurun_out/r02_c7_host_mem.txt; nproc >> gpurun_out/r02_c7_host_mem.txt
MEM_GB=$(awk '/MemTotal/ {print int($2/1048576)}' /proc/meminfo)
E2E=""; if [ "$MEM_GB" -lt 500 ]; then E2E="--no-e2e"; fi   # c5 holds ~20 GB of host arrays per rank (+ ~9 GB pinned for the e2e leg)
timeout 700 $TR --master-port 29553 bench.py --gpus 8 --workload c5 --steps 6 --warmup 3 $E2E > gpurun_out/r02_c7_bench_c5_n8.json 2> gpurun_out/r02_c7_bench_c5_n8.err
head -c 2500 gpurun_out/r02_c7_bench_c4_n8.json; echo; cat gpurun_out/r02_c7_suite_n8.jsonl; head -c 2500 gpurun_out/r02_c7_bench_c5_n8.json; echo; tail -5 gpurun_out/r02_c7_bench_c5_n8.err; cat gpurun_out/r02_c7_host_mem.txt
