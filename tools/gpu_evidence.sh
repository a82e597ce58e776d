#!/bin/bash
# One GPU-box session that produces the round's tracked evidence (tools/make_profiles.py <tag> afterwards):
# smoke, parity suite, bench (fp32 headline, fp64, reference arm), ncu launch list of the bench command,
# ncu --set full captures of the dominant kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
timeout 300 python bench.py --steps 40 --warmup 3 --dtype f64 --no-extras > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python bench.py --impl reference-cuda > gpurun_out/bench_refcuda.json 2> gpurun_out/bench_refcuda.err
LQPB_TC_FUSED=0 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-extras > gpurun_out/bench_f32_perphase.json 2> gpurun_out/bench_f32_perphase.err
timeout 300 python tools/unroll_bench.py > gpurun_out/unroll_bench.jsonl 2>&1
cat gpurun_out/bench_f32.json gpurun_out/bench_f64.json gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_f32.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'iterate_kernel' -s 3 -c 1 -o gpurun_out/iterate_f32 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_iter32.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'iterate_kernel' -s 3 -c 1 -o gpurun_out/iterate_f64 -f \
  python bench.py --steps 1 --warmup 3 --dtype f64 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_iter64.log 2>&1
# the fused block-sweep kernel (default at B = 128), then the per-phase kernels it replaced there (LQPB_TC_FUSED=0: still the
# path of B < 64, B > 148 and sweeps of more than 4 block rows)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_fused_kernel' -s 6 -c 2 -o gpurun_out/tc_fused_f32 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_fused32.log 2>&1
LQPB_TC_FUSED=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_tile_kernel' -s 24 -c 2 -o gpurun_out/tc_tile_f32 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_tile32.log 2>&1
LQPB_TC_FUSED=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_pivot8' -s 12 -c 1 -o gpurun_out/tc_pivot_f32 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_pivot32.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'d_tile_kernel|d_pivot8_kernel' -s 9 -c 3 -o gpurun_out/f64_block -f \
  python bench.py --steps 1 --warmup 3 --dtype f64 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_f64blk.log 2>&1
LQPB_FACTOR=gj timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gj_inverse_kernel' -s 6 -c 1 -o gpurun_out/gj_f64 -f \
  python bench.py --steps 1 --warmup 3 --dtype f64 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_gj64.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_f64.csv \
  python bench.py --dtype f64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_launch64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'iterate_kernel' -s 3 -c 1 -o gpurun_out/iterate_f32_dz1000 -f \
  python bench.py --dz 1000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_iter1000.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'iterate_row_kernel' -s 3 -c 1 -o gpurun_out/iterate_row_f32_dz100 -f \
  python bench.py --dz 100 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_row100.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'iterate_res_kernel' -s 3 -c 1 -o gpurun_out/iterate_res_f32_dz250 -f \
  python bench.py --dz 250 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_res250.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'unroll_reverse_kernel|tape_outer_kernel' -c 2 -o gpurun_out/unroll_f32 -f \
  python tools/unroll_bench.py --steps 1 > gpurun_out/ncu_unroll.log 2>&1
ls -la gpurun_out
