#!/bin/bash
# one `ncu --set full` capture of the scorer (after a plain run of the same command that exited 0): tools/gpu_ncu_score.sh TAG cfg events
mkdir -p gpurun_out
T=${1:-r02n}; C=${2:-cfg5}; E=${3:-32000}
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-others --workload $C --events $E"
$B > gpurun_out/${T}_plain.log 2>&1; echo "plain rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"score_tc_kernel" -s 1 -c 1 -f -o gpurun_out/${T}_${C}_score $B > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/${T}_${C}_score.ncu-rep
