#!/bin/bash
# tools/quick2.sh: GPU parity tests, then cfg2 bench under a few env settings (one line each)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/q2.err | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(j['value']/1e6,1), 'M motifs/s', round(j['ms_per_step'],3), 'ms', {k:round(v,3) for k,v in j['roofline']['stage_ms_per_step'].items()})
except Exception as e: print('$*', 'FAILED', e)
"; tail -n 3 gpurun_out/q2.err | grep -v "^$" | head -3; }
for cfg in "$@"; do run $cfg; done
