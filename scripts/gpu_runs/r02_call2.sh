#!/bin/bash
# Round 2, call 2 (1 GPU): new GPU tests (order, goodness, local stress, scale bands), tile-order / L2-residency experiment.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_order_pinned.py tests/test_goodness.py tests/test_gpu_scale.py -q -m gpu 2>&1 | tail -30 > gpurun_out/r02_c2_pytest_new.log
timeout 600 python scripts/gpu_exp_order.py c4 > gpurun_out/r02_c2_exp_order_c4.log 2>&1
timeout 300 python scripts/gpu_exp_order.py mid > gpurun_out/r02_c2_exp_order_mid.log 2>&1
timeout 300 python scripts/gpu_exp_order.py longthin > gpurun_out/r02_c2_exp_order_longthin.log 2>&1
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r02_c2_pytest_all.log
cat gpurun_out/r02_c2_pytest_new.log gpurun_out/r02_c2_exp_order_c4.log gpurun_out/r02_c2_exp_order_mid.log gpurun_out/r02_c2_exp_order_longthin.log; tail -5 gpurun_out/r02_c2_pytest_all.log
