#!/bin/bash
# round-2 first GPU pass: tests, default bench (cfg5), launch list + full captures for cfg5 and cfg4 (D = 172)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_gpu_tests.log
tail -5 gpurun_out/r02a_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench_cfg5.json 2> gpurun_out/r02a_bench_cfg5.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02a_bench_cfg5.json; tail -5 gpurun_out/r02a_bench_cfg5.err
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-others"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_cfg5.csv $B --events 32000 > gpurun_out/r02a_ncu_l5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"score_tc_kernel|sample_walks_kernel" -s 2 -c 2 -f -o gpurun_out/r02a_cfg5 $B --events 16000 > gpurun_out/r02a_ncu_f5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"score_tc_kernel" -s 1 -c 1 -f -o gpurun_out/r02a_cfg4 $B --workload cfg4 --events 4000 > gpurun_out/r02a_ncu_f4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
