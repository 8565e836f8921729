#!/bin/bash
# Round-2 evidence at the FINAL state (walk-group scorer): the default bench line (what the driver runs), the reference arm, the ncu launch
# list and one `ncu --set full` capture of every per-step kernel on cfg5 plus the scorer on cfg4 (node_dim 172).  Every ncu pass follows a
# plain run of the same command that exited 0.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_cfg5.json 2> gpurun_out/r02b_bench_cfg5.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err; echo "ref rc=$?"
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-others --events 32000"
$B > gpurun_out/r02b_plain.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_cfg5.csv $B > gpurun_out/r02b_ncu_l.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"score_tc_kernel|sample_walks_kernel|edge_identity_kernel|sample_hop_kernel|time_std_kernel" -s 5 -c 5 -f -o gpurun_out/r02b_cfg5 $B > gpurun_out/r02b_ncu_f5.log 2>&1
B4="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-others --workload cfg4 --events 4000"
$B4 > gpurun_out/r02b_plain4.log 2>&1; echo "plain4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"score_tc_kernel" -s 1 -c 1 -f -o gpurun_out/r02b_cfg4 $B4 > gpurun_out/r02b_ncu_f4.log 2>&1
ls -la gpurun_out/r02b_*.ncu-rep
python - <<'P'
import json
for f in ("gpurun_out/r02b_bench_cfg5.json","gpurun_out/r02b_bench_reference.json"):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(j['value']/1e6,3),'M/s', 'e2e', round(j['e2e']['value']/1e6,3), j.get('clocks'))
    except Exception as e: print(f,'ERR',e)
P
