#!/bin/bash
# A/B of build-time variants: tools/gpu_ab_defs.sh TAG "cfgs" "DEFS1" "DEFS2" ...   ("-" = no extra defines)
mkdir -p gpurun_out
T=$1; CFGS=$2; shift 2
Q="--no-cpu-baseline --no-others --no-e2e --steps 5 --warmup 3"
i=0
for defs in "$@"; do
  if [ "$defs" = "-" ]; then unset TEMPME_BUILD_DEFS; else export TEMPME_BUILD_DEFS="$defs"; fi
  python -c "from tempme_b200 import build as b; b.build()" > gpurun_out/${T}_build$i.log 2>&1 || { echo "build failed for $defs"; tail -5 gpurun_out/${T}_build$i.log; continue; }
  if [ $i = 0 ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "walk_group or encoder_vs_oracle or encoder_golden" > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_tests.log; fi
  for c in $CFGS; do
    timeout 600 python bench.py $Q --workload $c > gpurun_out/${T}_${c}_v$i.json 2> gpurun_out/${T}_${c}_v$i.err
    python - gpurun_out/${T}_${c}_v$i.json "$defs" <<'P'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j['roofline']
    print(sys.argv[2], sys.argv[1], round(j['value']/1e6,1),'M/s', {k:round(v,2) for k,v in r['stage_ms_per_step'].items()})
except Exception as e: print(sys.argv[1],'ERR',e)
P
  done
  i=$((i+1))
done
