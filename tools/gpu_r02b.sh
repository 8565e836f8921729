#!/bin/bash
# A/B: hash edge identity + L2 discard of the h scratch (cfg5), 512-thread tiles at D = 172 (cfg4)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_gpu_tests.log
tail -3 gpurun_out/r02b_gpu_tests.log
Q="--no-cpu-baseline --no-others --no-e2e --steps 5 --warmup 3"
python bench.py $Q > gpurun_out/r02b_cfg5_discard.json 2> gpurun_out/r02b.err
TEMPME_TC_NO_DISCARD=1 python bench.py $Q > gpurun_out/r02b_cfg5_nodiscard.json 2>> gpurun_out/r02b.err
python bench.py $Q --workload cfg4 > gpurun_out/r02b_cfg4_cw16.json 2>> gpurun_out/r02b.err
TEMPME_TC_CW=8 python bench.py $Q --workload cfg4 > gpurun_out/r02b_cfg4_cw8.json 2>> gpurun_out/r02b.err
TEMPME_TC_CW=8 python bench.py $Q --workload cfg3 > gpurun_out/r02b_cfg3_cw8.json 2>> gpurun_out/r02b.err
python bench.py $Q --workload cfg5 --chunk 8000 > gpurun_out/r02b_cfg5_chunk8k.json 2>> gpurun_out/r02b.err
python bench.py $Q --workload cfg5 --chunk 32000 > gpurun_out/r02b_cfg5_chunk32k.json 2>> gpurun_out/r02b.err
for f in gpurun_out/r02b_cfg*.json; do python - "$f" <<'P'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j['roofline']
    print(sys.argv[1], round(j['value']/1e6,1),'M/s', {k:round(v,2) for k,v in r['stage_ms_per_step'].items()})
except Exception as e: print(sys.argv[1],'ERR',e)
P
done
tail -5 gpurun_out/r02b.err
