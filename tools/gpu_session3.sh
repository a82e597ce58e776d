#!/bin/bash
# GPU session: parity suite, bench f32/f64 (+ reference arm optionally)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err; cat gpurun_out/bench_f32.json; tail -3 gpurun_out/bench_f32.err
timeout 300 python bench.py --dtype f64 --no-cpu-baseline > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err; cat gpurun_out/bench_f64.json; tail -3 gpurun_out/bench_f64.err
