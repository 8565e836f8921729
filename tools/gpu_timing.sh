#!/bin/bash
# per-round clock stamps of CTA 0 (diagnostic build with -DTM_TC_TIMING), walk-group mode vs per-walk evaluation
mkdir -p gpurun_out
T=${1:-r02t}
export TEMPME_BUILD_TIMING=1
python -c "from tempme_b200 import build as b; b.build()" > gpurun_out/${T}_build.log 2>&1; echo "build rc=$?"
Q="--no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 1 --events 32000"
for c in ${2:-cfg5}; do
  TEMPME_TC_TIMING=1 TEMPME_TC_DEBUG=1 timeout 600 python bench.py $Q --workload $c > gpurun_out/${T}_${c}_share.json 2> gpurun_out/${T}_${c}_share.err
  TEMPME_TC_NO_SHARE=1 TEMPME_TC_TIMING=1 TEMPME_TC_DEBUG=1 timeout 600 python bench.py $Q --workload $c > gpurun_out/${T}_${c}_plain.json 2> gpurun_out/${T}_${c}_plain.err
done
grep -c "tc timing" gpurun_out/${T}_*.err
