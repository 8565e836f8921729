#!/bin/bash
# A/B of build-time variants with the scorer's DRAM bytes (ncu metrics pass after the plain run): tools/gpu_ab_dram.sh TAG "cfgs" "DEFS1" ...
mkdir -p gpurun_out
T=$1; CFGS=$2; shift 2
Q="--no-cpu-baseline --no-others --no-e2e --steps 5 --warmup 3"
i=0
for defs in "$@"; do
  if [ "$defs" = "-" ]; then unset TEMPME_BUILD_DEFS; else export TEMPME_BUILD_DEFS="$defs"; fi
  python -c "from tempme_b200 import build as b; b.build()" > gpurun_out/${T}_build$i.log 2>&1 || { echo "build failed for $defs"; tail -5 gpurun_out/${T}_build$i.log; continue; }
  if [ $i = 0 ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "walk_group or encoder_vs_oracle or encoder_golden" > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_tests.log; fi
  for c in $CFGS; do
    timeout 600 python bench.py $Q --workload $c > gpurun_out/${T}_${c}_v$i.json 2> gpurun_out/${T}_${c}_v$i.err
    python - gpurun_out/${T}_${c}_v$i.json "$defs" <<'P'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j['roofline']
    print(sys.argv[2], sys.argv[1], round(j['value']/1e6,1),'M/s', {k:round(v,2) for k,v in r['stage_ms_per_step'].items()})
except Exception as e: print(sys.argv[1],'ERR',e)
P
    EV=32000; [ $c = cfg4 ] && EV=4000; [ $c = cfg3 ] && EV=4000; [ $c = cfg1 ] && EV=2000
    timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"score_tc_kernel|sample_walks" -s 2 -c 2 --csv --log-file gpurun_out/${T}_${c}_v${i}_dram.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-others --workload $c --events $EV > /dev/null 2>&1
    python - gpurun_out/${T}_${c}_v${i}_dram.csv <<'P'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print('   ', d['Kernel Name'][:28], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
P
  done
  i=$((i+1))
done
