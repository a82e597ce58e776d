#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py > gpurun_out/check.txt 2>&1; echo "check rc=$?" >> gpurun_out/check.txt
cat gpurun_out/check.txt
timeout 200 python tools/iter_tune.py f32 > gpurun_out/tune_f32.txt 2>&1; cat gpurun_out/tune_f32.txt
timeout 200 python tools/iter_tune.py f64 > gpurun_out/tune_f64.txt 2>&1; cat gpurun_out/tune_f64.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
