#!/bin/bash
# tools/cfgs.sh: one short bench line per workload (no CPU baseline, no e2e)
for w in "$@"; do
python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/cfgs.err | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(j['value']/1e6,1), 'M motifs/s', round(j['ms_per_step'],3), 'ms', {k:round(v,3) for k,v in j['roofline']['stage_ms_per_step'].items()})
except Exception as e: print('$w', 'FAILED', e)
"; done
