#!/bin/bash
# tools/allcfg.sh: one bench line per workload (no CPU baseline), summary per line
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/all_$w.json 2> gpurun_out/all_$w.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/all_$w.json"))
    r = j["roofline"]
    print("$w", round(j["value"] / 1e6, 1), "M motifs/s", round(j["ms_per_step"], 3), "ms  e2e", round(j["e2e"]["value"] / 1e6, 1), "| top:", r["kernel"][:24], "frac", round(r["frac"], 4), {k: round(v, 3) for k, v in r["stage_ms_per_step"].items()}, j["clocks"])
except Exception as e:
    print("$w", "no result:", e, open("gpurun_out/all_$w.err").read()[-400:])
PY
done
