#!/bin/bash
# GPU test suite + smoke + a short default bench line
mkdir -p gpurun_out
T=${1:-r02}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -25 gpurun_out/${T}_gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 --no-others > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - gpurun_out/${T}_bench.json <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j['roofline']
print(round(j['value']/1e6,1),'M/s e2e',round(j['e2e']['value']/1e6,1), {k:round(v,2) for k,v in r['stage_ms_per_step'].items()}, 'frac',round(r['frac'],4), j['e2e']['api'])
P
tail -3 gpurun_out/${T}_bench.err
