#!/bin/bash
# Short evidence session (kernels unchanged since the last full gpu_evidence.sh run): smoke, parity suite, the default bench
# line, the ncu launch list of the bench command, the end-to-end timeline.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
tail -2 gpurun_out/bench_f32.err
timeout 300 python tools/e2e_trace.py > gpurun_out/e2e_trace.txt 2>&1
cat gpurun_out/e2e_trace.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_f32.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_launch.log 2>&1
timeout 300 python bench.py --batch 32 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/bench_f32_b32.json 2> gpurun_out/bench_f32_b32.err
LQPB_ITER_SPLIT=0 timeout 300 python bench.py --batch 32 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/bench_f32_b32_unsplit.json 2> gpurun_out/bench_f32_b32_unsplit.err
timeout 300 python tools/exp2_breakdown.py > gpurun_out/exp2_breakdown.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'iterate_split_kernel' -s 3 -c 1 -o gpurun_out/iterate_split_f32 -f \
  python bench.py --batch 32 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_split32.log 2>&1
ls -la gpurun_out | tail -12
