#!/bin/bash
# sweep of the block-aggregation tuning knobs on config 2 (resident only): "ENV=VAL ENV=VAL" per line
while read -r cfg; do
  [ -z "$cfg" ] && continue
  env $cfg timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/sw.json 2> gpurun_out/sw.err || tail -3 gpurun_out/sw.err
  python - "$cfg" <<PY
import json, sys
d = json.load(open("gpurun_out/sw.json"))
r = d["stages"]["raw_ms"]
print("%-50s %.2f ms/step  kA %.2f  kB %.2f  k4 %.2f" % (sys.argv[1], d["ms_per_step"], r["ms_sort"], r["ms_count"], r["ms_dedup"]))
PY
done
