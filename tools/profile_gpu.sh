#!/bin/bash
# ncu evidence for the round (run under gpurun, 1 GPU): launch list of a short bench run and a full
# capture of the dominant kernel.  Usage: tools/profile_gpu.sh <tag> [dtype]
TAG=${1:-r01}; DT=${2:-f32}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}_${DT}.csv \
    python bench.py --steps 2 --warmup 3 --dtype $DT --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu_${TAG}_${DT}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:iterate_kernel -s 3 -c 1 -f -o gpurun_out/prof_iterate_${TAG}_${DT} \
    python bench.py --steps 2 --warmup 3 --dtype $DT --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_${TAG}_${DT}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gj_inverse -s 3 -c 1 -f -o gpurun_out/prof_gj_${TAG}_${DT} \
    python bench.py --steps 2 --warmup 3 --dtype $DT --no-cpu-baseline --no-e2e > gpurun_out/ncu_gj_${TAG}_${DT}.log 2>&1
ls -la gpurun_out
