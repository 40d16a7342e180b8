#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python scripts/gpu_exp_policy.py c4 > gpurun_out/r02_c9_exp_policy_c4.log 2>&1
timeout 600 python -m pytest tests/test_gpu_scale.py -q -m gpu 2>&1 | tail -30 > gpurun_out/r02_c9_pytest_scale.log
cat gpurun_out/r02_c9_exp_policy_c4.log; tail -25 gpurun_out/r02_c9_pytest_scale.log
