#!/bin/bash
for c in 4 3 2; do
  TEMPME_TC_EVENT_CTAS=$c TEMPME_TC_DEBUG=1 TEMPME_TC_SERIAL=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/ev$c.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('event ctas=$c serial', round(j['value']/1e6,1), {k:round(v,3) for k,v in j['roofline']['stage_ms_per_step'].items()})"
  grep -m1 "^\[tc\]" gpurun_out/ev$c.err
  TEMPME_TC_EVENT_CTAS=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('event ctas=$c 2-stream', round(j['value']/1e6,1), round(j['ms_per_step'],3))"
done
