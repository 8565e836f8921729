#!/bin/bash
# walk sampler: threads per block A/B (TEMPME_WALKS_BLOCK), parity tests under each value first
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-others --no-e2e --steps 5 --warmup 3"
for wb in 256 128 64; do
  export TEMPME_WALKS_BLOCK=$wb
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x -k "oracle_parity or golden_csr or pipeline or bench_graph or injected or shard_offset" 2>&1 | tail -1
  for c in cfg5 cfg4 cfg2; do
    python bench.py $Q --workload $c 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print('wb=$wb $c', round(j['value']/1e6,1),'M/s', {k:round(v,2) for k,v in r['stage_ms_per_step'].items()})"
  done
done
