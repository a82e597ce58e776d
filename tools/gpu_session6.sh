#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py 512 16 2>&1 | grep "tc:\|x  \|dQ\|SPD\|KKT"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_f32.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()})
PY
LQPB_TC_GROUPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tc_ -s 26 -c 26 python tools/tc_prof.py 500 128 3 2>&1 | grep -A1 "tc_.*(" | grep -v "^--" | paste - - | awk '{print $1, $(NF)}'
