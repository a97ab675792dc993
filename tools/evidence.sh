#!/bin/bash
# bench lines + launch list + ncu summaries that profiles/ cites (run under gpurun, one GPU)
TAG=${1:-r1j}
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
python bench.py --clusters 200 --targets-all --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg3.json 2> gpurun_out/c3.err
python bench.py --samples 10000 --clusters 96 --no-cpu-baseline --steps 3 > gpurun_out/bench_${TAG}_cfg4.json 2> gpurun_out/c4.err
python bench.py --samples 50000 --clusters 16 --consider-missing --no-cpu-baseline --steps 3 > gpurun_out/bench_${TAG}_cfg5.json 2> gpurun_out/c5.err
bash tools/launch_list.sh ${TAG} --clusters 200 > gpurun_out/ll_${TAG}.txt 2>&1
bash tools/ncu_block.sh ${TAG} > gpurun_out/ncu_${TAG}.txt 2>&1
python profiles/ncu_summary.py gpurun_out/raw_${TAG}.csv > gpurun_out/ncu_full_summary_${TAG}.txt 2>&1
tail -c 600 gpurun_out/bench_${TAG}.err
for f in gpurun_out/bench_${TAG}*.json; do python - "$f" <<PY
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    e = d.get("e2e") or {}
    print(sys.argv[1].split("/")[-1], "value %.3g" % d.get("value", 0), "ms/step %.2f" % d.get("ms_per_step", 0), "e2e %.3g" % e.get("value", 0), d.get("engine"), (d.get("cpu_baseline") or {}).get("value"))
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
