#!/bin/bash
# A/B of a run-time knob: tools/gpu_ab_env.sh TAG "cfgs" VAR   (VAR unset vs VAR=1), walk-group tests first
mkdir -p gpurun_out
T=$1; CFGS=$2; VAR=$3
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "walk_group or encoder_vs_oracle or encoder_golden or enhance" > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_tests.log
Q="--no-cpu-baseline --no-others --no-e2e --steps 5 --warmup 3"
for c in $CFGS; do
  for v in 0 1; do
    if [ $v = 1 ]; then export $VAR=1; else unset $VAR; fi
    python bench.py $Q --workload $c 2>gpurun_out/${T}_${c}_$v.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print('$VAR=$v $c', round(j['value']/1e6,1),'M/s', {k:round(v,2) for k,v in r['stage_ms_per_step'].items()})"
  done
done
