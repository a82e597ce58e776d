for v in 1 0; do
LQPB_HOST_FULL_Q=$v timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); print('full_q', os.environ.get('LQPB_HOST_FULL_Q'), round(d['value']), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3))"
done
