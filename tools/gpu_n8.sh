#!/bin/bash
# 8-GPU check of the default bench under torchrun (what the driver launches for N = 8), short
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r03_bench_n$N.json 2> gpurun_out/r03_bench_n$N.err; echo "bench rc=$?"
python - gpurun_out/r03_bench_n$N.json <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j['roofline']
print(j['n_gpus'],'GPUs', round(j['value']/1e6,1),'M/s e2e',round(j['e2e']['value']/1e6,1), {k:round(v,2) for k,v in r['stage_ms_per_step'].items()}, j['setup'], j['multi_gpu_check'], j['config']['exchange'])
P
