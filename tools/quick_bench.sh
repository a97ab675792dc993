#!/bin/bash
# quick resident-only bench of config 2 (no e2e, no cpu baseline) + stage split
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/qb.json 2> gpurun_out/qb.err || tail -5 gpurun_out/qb.err
python - <<PY
import json
d = json.load(open("gpurun_out/qb.json"))
print("value %.2f Gb/s  ms/step %.2f" % (d["value"]/1e9, d["ms_per_step"]))
print(d["stages"]["raw_ms"])
print("U", d["unique_kmers_per_step_per_gpu"], "rows", d["rows_per_step_per_gpu"], "pat", d["patterns_per_gpu"])
PY
