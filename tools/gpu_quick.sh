#!/bin/bash
# quick check: sampler parity tests + one short cfg5 line (and any extra workloads given as arguments)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x -k "oracle_parity or bench_graph or golden_csr or pipeline or device_build" 2>&1 | tail -3
Q="--no-cpu-baseline --no-others --no-e2e --steps 5 --warmup 3"
for c in cfg5 "$@"; do python bench.py $Q --workload $c 2>> gpurun_out/quick.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print('$c', round(j['value']/1e6,1),'M/s', {k:round(v,2) for k,v in r['stage_ms_per_step'].items()})"; done
tail -2 gpurun_out/quick.err
