set -x
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --events 16000"
$CMD > gpurun_out/r01e_plain.json 2> gpurun_out/r01e_plain.err; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches.csv $CMD > gpurun_out/r01e_ncu1.log 2>&1; echo "launch-list rc=$?"
$CMD > /dev/null 2>&1; echo "plain2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'score_tc_kernel|sample_walks_kernel' -s 4 -c 2 -o gpurun_out/prof_r01e -f $CMD > gpurun_out/r01e_ncu2.log 2>&1; echo "full rc=$?"
python bench.py > gpurun_out/r01e_bench_cfg2.json 2> gpurun_out/r01e_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r01e_bench_cfg2.json
