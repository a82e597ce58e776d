#!/bin/bash
# A/B of the multi-stream slices of the tensor-core factorisation + ncu capture of the tile kernels
mkdir -p gpurun_out
for G in 1 2 3 4; do
  echo "== LQPB_TC_GROUPS=$G"
  LQPB_TC_GROUPS=$G timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 30 > gpurun_out/bench_g$G.json 2> gpurun_out/bench_g$G.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_g$G.json"))
print("G=$G value", round(d["value"]), "ms", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()})
PY
done
LQPB_TC_GROUPS=2 timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_tile_kernel -s 24 -c 2 -f -o gpurun_out/tc_tile_persist python tools/tc_prof.py 500 128 3 > gpurun_out/ncu_tile.log 2>&1
tail -3 gpurun_out/ncu_tile.log
timeout 120 python tools/e2e_breakdown.py f32 > gpurun_out/e2e_breakdown.txt 2>&1; cat gpurun_out/e2e_breakdown.txt
