#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02k_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_gpu_tests.log
tail -4 gpurun_out/r02k_gpu_tests.log
Q="--no-cpu-baseline --no-others --no-e2e --steps 5 --warmup 3"
for c in cfg4 cfg3 cfg1; do python bench.py $Q --workload $c > gpurun_out/r02k_$c.json 2>> gpurun_out/r02k.err; done
for f in gpurun_out/r02k_cfg*.json; do python - "$f" <<'P'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j['roofline']
    print(sys.argv[1], round(j['value']/1e6,1),'M/s', {k:round(v,2) for k,v in r['stage_ms_per_step'].items()})
except Exception as e: print(sys.argv[1],'ERR',e)
P
done
tail -3 gpurun_out/r02k.err
