#!/bin/bash
# sweep of the block-aggregation tuning knobs on config 2 (resident only)
for cfg in "16 512" "16 1024" "32 1024" "32 2048" "64 2048"; do
  set -- $cfg
  PF_BLOCK_WINDOWS=$1 PF_BLOCK_SLOTS=$2 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/sw.json 2> gpurun_out/sw.err || tail -3 gpurun_out/sw.err
  python - <<PY
import json
d = json.load(open("gpurun_out/sw.json"))
r = d["stages"]["raw_ms"]
print("B=$1 slots=$2: %.2f ms/step  kA %.2f  kB %.2f  k4 %.2f" % (d["ms_per_step"], r["ms_sort"], r["ms_count"], r["ms_dedup"]))
PY
done
