#!/bin/bash
# per-launch device times of one full config-2 step (cold cache, serialised: compare SHARES)
TAG=${1:-ll}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra ${@:2} > gpurun_out/ll_${TAG}.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_${TAG}.csv")) if len(r) > 10 and r[0].isdigit()]
# keep the last step: launches after the last pf "k4_commit" of the previous step are hard to split; show totals / 2
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][:60]
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += float(r[-1].replace(",", "")) / 1e3
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("%-62s n=%3d  total %9.1f us" % (k, n, us))
PY
