#!/bin/bash
# tools/quick.sh [tag]: GPU parity tests + cfg2 bench (serial and two-stream), one summary line each
tag=${1:-q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
TEMPME_TC_SERIAL=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_serial.json 2>gpurun_out/${tag}_serial.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}.json 2>gpurun_out/${tag}.err
tail -n 2 gpurun_out/${tag}.err
python - <<PY
import json
for f in ("${tag}_serial", "${tag}"):
    try:
        j = json.load(open("gpurun_out/%s.json" % f))
        print(f, round(j["value"] / 1e6, 1), "M motifs/s", round(j["ms_per_step"], 3), "ms", {k: round(v, 3) for k, v in j["roofline"]["stage_ms_per_step"].items()}, j["roofline"].get("kernel_ms_concurrent"))
    except Exception as e:
        print(f, "no result:", e)
PY
