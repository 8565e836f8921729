#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02m_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_gpu_tests.log
grep -E "^FAILED|passed|failed" gpurun_out/r02m_gpu_tests.log | tail -12
Q="--no-cpu-baseline --no-others --no-e2e --steps 5 --warmup 3"
for c in cfg5 cfg4 cfg3 cfg1 cfg2; do
  TEMPME_TC_DEBUG=1 python bench.py $Q --workload $c > gpurun_out/r02m_$c.json 2> gpurun_out/r02m_$c.err
  TEMPME_EDGE_PROJECTION=0 python bench.py $Q --workload $c > gpurun_out/r02m_${c}_noproj.json 2>> gpurun_out/r02m.err
  grep "\[tc\]" gpurun_out/r02m_$c.err | sort | uniq -c | head -2
done
for f in gpurun_out/r02m_cfg*.json; do python - "$f" <<'P'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j['roofline']
    print(sys.argv[1], round(j['value']/1e6,1),'M/s', {k:round(v,2) for k,v in r['stage_ms_per_step'].items()})
except Exception as e: print(sys.argv[1],'ERR',e)
P
done
tail -3 gpurun_out/r02m.err
