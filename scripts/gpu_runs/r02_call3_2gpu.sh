#!/bin/bash
# Round 2, call 3 (2 GPUs): the multi-GPU tests that a 1-GPU box skips, then c4 at N=2 in every mode and the rank-locally
# generated path-sharded stand-in (c5s).
#   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_runs/r02_call3_2gpu.sh'
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02_c3_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_host.py tests/test_gpu_more.py -q -k "two_rank or single_process or two_gpus or sharded" 2>&1 | tail -40 > gpurun_out/r02_c3_pytest_2gpu.log
for MODE in auto hybrid peer; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --multi $MODE \
      > gpurun_out/r02_c3_bench_c4_n2_$MODE.json 2> gpurun_out/r02_c3_bench_c4_n2_$MODE.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --workload c5s \
      > gpurun_out/r02_c3_bench_c5s_n2.json 2> gpurun_out/r02_c3_bench_c5s_n2.err
tail -15 gpurun_out/r02_c3_pytest_2gpu.log
for f in gpurun_out/r02_c3_bench_*.json; do echo $f; head -c 1500 $f; echo; tail -3 ${f%.json}.err; done
