#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02f_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_gpu_tests.log
tail -6 gpurun_out/r02f_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-others --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
python - gpurun_out/r02f_bench.json <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j['roofline']
print(round(j['value']/1e6,1),'M/s e2e',round(j['e2e']['value']/1e6,1), {k:round(v,2) for k,v in r['stage_ms_per_step'].items()}, j['setup'])
P
tail -3 gpurun_out/r02f_bench.err
