#!/bin/bash
# 2-GPU validation: multi-GPU test, then the default bench under torchrun (what the driver launches for N > 1)
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02g_multi_test.log 2>&1; echo "multi test rc=$?"; tail -3 gpurun_out/r02g_multi_test.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02g_bench_n$N.json 2> gpurun_out/r02g_bench_n$N.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02g_bench_n$N.err
python - gpurun_out/r02g_bench_n$N.json <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j['roofline']
print(j['n_gpus'],'GPUs', round(j['value']/1e6,1),'M/s e2e',round(j['e2e']['value']/1e6,1), {k:round(v,2) for k,v in r['stage_ms_per_step'].items()}, j['setup'], j['multi_gpu_check'], j['config']['exchange'])
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02g_ref_n$N.json 2> gpurun_out/r02g_ref_n$N.err; echo "ref rc=$?"
cut -c1-400 gpurun_out/r02g_ref_n$N.json
