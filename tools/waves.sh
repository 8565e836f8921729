#!/bin/bash
for w in 1 2 4 8 16; do
  TEMPME_TC_SLAB_WAVES=$w TEMPME_TC_SERIAL=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('waves=$w serial', round(j['value']/1e6,1), {k:round(v,3) for k,v in j['roofline']['stage_ms_per_step'].items()})"
  TEMPME_TC_SLAB_WAVES=$w python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('waves=$w 2-stream', round(j['value']/1e6,1), round(j['ms_per_step'],3))"
done
