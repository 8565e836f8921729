#!/bin/bash
# A/B of the scorer variants on one box: tools/ab.sh  (writes gpurun_out/ab_*.json)
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err
TEMPME_TC_MOTIF=ts $B > gpurun_out/ab_mts.json 2> gpurun_out/ab_mts.err
TEMPME_TC_EVENT=ts $B > gpurun_out/ab_ets.json 2> gpurun_out/ab_ets.err
TEMPME_TC_EVENT=ts TEMPME_TC_MOTIF=ts $B > gpurun_out/ab_both.json 2> gpurun_out/ab_both.err
TEMPME_TC_SERIAL=1 $B > gpurun_out/ab_default_serial.json 2> gpurun_out/ab_default_serial.err
TEMPME_TC_SERIAL=1 TEMPME_TC_EVENT=ts TEMPME_TC_MOTIF=ts $B > gpurun_out/ab_both_serial.json 2> gpurun_out/ab_both_serial.err
for f in gpurun_out/ab_*.json; do echo "$f: $(python -c "
import json,sys
j=json.load(open('$f'))
print(round(j['value']/1e6,1), round(j['ms_per_step'],3), {k:round(v,3) for k,v in j['roofline']['stage_ms_per_step'].items()}, j['roofline'].get('kernel_ms_concurrent'))
" 2>&1 | tail -1)"; done
