#!/bin/bash
# Round 2, first single-GPU call: everything that was written after round 1's GPU budget was spent.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_runs/r02_first_1gpu.sh'
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
make -C odgi_b200/host > /dev/null 2>&1
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err
timeout 300 python scripts/gpu_exp_1d_full.py mid > gpurun_out/r02_exp_1d_full_mid.log 2>&1
timeout 200 python scripts/gpu_exp_hub_onset.py > gpurun_out/r02_exp_hub_onset.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_smoke.log | tail -2; head -c 600 gpurun_out/r02_bench_c4_n1.json; echo; cat gpurun_out/r02_exp_1d_full_mid.log; cat gpurun_out/r02_exp_hub_onset.log
