#!/bin/bash
# L2 set-aside for the scorer's scratch: A/B over sizes, walk-group mode and per-walk evaluation
mkdir -p gpurun_out
T=${1:-r02p}
Q="--no-cpu-baseline --no-others --no-e2e --steps 5 --warmup 3"
for c in ${2:-cfg5}; do
 for mb in ${3:-0 32 64 96}; do
  for v in 0 1; do
    if [ $v = 1 ]; then export TEMPME_TC_NO_SHARE=1; else unset TEMPME_TC_NO_SHARE; fi
    TEMPME_TC_DEBUG=1 TEMPME_TC_L2_PERSIST_MB=$mb timeout 600 python bench.py $Q --workload $c > gpurun_out/${T}_${c}_p${mb}_noshare$v.json 2> gpurun_out/${T}_${c}_p${mb}_noshare$v.err
    python - gpurun_out/${T}_${c}_p${mb}_noshare$v.json <<'P'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j['roofline']
    print(sys.argv[1], round(j['value']/1e6,1),'M/s', {k:round(v,2) for k,v in r['stage_ms_per_step'].items()})
except Exception as e: print(sys.argv[1],'ERR',e)
P
  done
 done
done
grep -h "L2 set-aside" gpurun_out/${T}_*.err | sort | uniq -c
