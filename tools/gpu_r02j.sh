#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02j_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_gpu_tests.log
tail -4 gpurun_out/r02j_gpu_tests.log
Q="--no-cpu-baseline --no-others --no-e2e --steps 5 --warmup 3"
python bench.py $Q > gpurun_out/r02j_cfg5.json 2> gpurun_out/r02j.err
TEMPME_L2_FETCH=32 python bench.py $Q > gpurun_out/r02j_cfg5_l2f32.json 2>> gpurun_out/r02j.err
for c in cfg4 cfg3 cfg1 cfg2; do python bench.py $Q --workload $c > gpurun_out/r02j_$c.json 2>> gpurun_out/r02j.err; done
for f in gpurun_out/r02j_cfg*.json; do python - "$f" <<'P'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j['roofline']
    print(sys.argv[1], round(j['value']/1e6,1),'M/s', {k:round(v,2) for k,v in r['stage_ms_per_step'].items()})
except Exception as e: print(sys.argv[1],'ERR',e)
P
done
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-others"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sample_walks_kernel" -s 1 -c 1 -f -o gpurun_out/r02j_walks $B --events 32000 > gpurun_out/r02j_ncu.log 2>&1
tail -3 gpurun_out/r02j.err
