#!/bin/bash
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for R in 1 8; do for M in ss ts; do
 TEMPME_TC_REPLICAS=$R TEMPME_TC_SERIAL=1 TEMPME_TC_MOTIF=$M $B 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('R=$R motif=$M', round(j['value']/1e6,1), {k:round(v,3) for k,v in j['roofline']['stage_ms_per_step'].items()})"
done; done
TEMPME_TC_REPLICAS=8 TEMPME_TC_TIMING=1 TEMPME_TC_MOTIF=ts python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --events 4000 2>&1 | grep -A 30 "TS kernel" | head -34
TEMPME_TC_REPLICAS=1 TEMPME_TC_TIMING=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --events 4000 2>&1 | grep -A 8 "event kernel" | head -12
