#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gpu_exp_quality.py mid > gpurun_out/exp_quality_mid.log 2>&1; cat gpurun_out/exp_quality_mid.log
timeout 900 python scripts/gpu_exp_quality.py c4 > gpurun_out/exp_quality_c4.log 2>&1; cat gpurun_out/exp_quality_c4.log
