#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py 512 16 2>&1 | tail -12
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for PV in 1 0; do
  LQPB_TC_PIVOT=$PV timeout 300 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/bench_pv$PV.json 2> gpurun_out/bench_pv$PV.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_pv$PV.json"))
print("PIVOT_V1=$PV value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()})
PY
done
