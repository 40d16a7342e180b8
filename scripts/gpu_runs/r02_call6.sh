#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python scripts/gpu_exp_depth.py > gpurun_out/r02_c6_exp_depth.log 2>&1
timeout 300 python -m pytest tests/test_gpu_lay.py tests/test_goodness.py tests/test_order_pinned.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_c6_pytest_lay.log
cat gpurun_out/r02_c6_exp_depth.log; tail -8 gpurun_out/r02_c6_pytest_lay.log
