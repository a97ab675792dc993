#!/bin/bash
# A/B of differently built libraries (PF_LIB_PATH) on the resident config-2 step: prints ms/step and the kA stage time
for lib in "$@"; do
  PF_LIB_PATH=$lib timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - "$lib" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/ab.json"))
    print(sys.argv[1].split("/")[-1], "ms/step %.3f" % d["ms_per_step"], d["stages"]["raw_ms"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
