#!/bin/bash
# compute-sanitizer over one forward + backward of every kernel family (GPU box).  Summaries -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for c in ${CASES:-rows packed stream fused adapt unroll f64 f64blk}; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_cases.py $c > gpurun_out/sanitizer_${tool}_$c.log 2>&1
    echo "$tool $c rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok ' gpurun_out/sanitizer_${tool}_$c.log | tr '\n' ' ')"
  done
done
