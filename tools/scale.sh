#!/bin/bash
# tools/scale.sh N...: the driver's scaling run (one bench line per N) on one box
for n in "$@"; do
  if [ "$n" = 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/scale_$n.err | tail -1 > gpurun_out/scale_$n.json
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/scale_$n.err | tail -1 > gpurun_out/scale_$n.json; fi
  python -c "
import json; j=json.load(open('gpurun_out/scale_$n.json')); print(j['n_gpus'], round(j['value']/1e6,1), 'M motifs/s', round(j['ms_per_step'],3), 'ms  e2e', round(j['e2e']['value']/1e6,1), {k:round(v,3) for k,v in j['roofline']['stage_ms_per_step'].items()})"
done
