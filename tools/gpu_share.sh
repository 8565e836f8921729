#!/bin/bash
# walk-group mode: its parity tests, then A/B bench lines (TEMPME_TC_NO_SHARE=1 = per-walk evaluation)
mkdir -p gpurun_out
T=${1:-r02s}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "walk_group or encoder_vs_oracle or pipeline or encoder_golden or enhance" > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${T}_tests.log
Q="--no-cpu-baseline --no-others --no-e2e --steps 5 --warmup 3"
for c in cfg5 cfg4 cfg3 cfg1; do
  for v in 0 1; do
    if [ $v = 1 ]; then export TEMPME_TC_NO_SHARE=1; else unset TEMPME_TC_NO_SHARE; fi
    timeout 600 python bench.py $Q --workload $c > gpurun_out/${T}_${c}_noshare$v.json 2> gpurun_out/${T}_${c}_noshare$v.err
    python - gpurun_out/${T}_${c}_noshare$v.json <<'P'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j['roofline']
    print(sys.argv[1], round(j['value']/1e6,1),'M/s', {k:round(v,2) for k,v in r['stage_ms_per_step'].items()})
except Exception as e: print(sys.argv[1],'ERR',e)
P
  done
done
unset TEMPME_TC_NO_SHARE
tail -3 gpurun_out/${T}_cfg5_noshare0.err
