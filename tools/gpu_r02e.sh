#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02e_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_gpu_tests.log
tail -6 gpurun_out/r02e_gpu_tests.log
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-others"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"score_tc_kernel" -s 1 -c 1 -f -o gpurun_out/r02e_cfg4_drain $B --workload cfg4 --events 4000 > gpurun_out/r02e_ncu_f4.log 2>&1
ls -la gpurun_out/r02e_cfg4_drain.ncu-rep
