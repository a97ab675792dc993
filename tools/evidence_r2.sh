#!/bin/bash
# Round-2 evidence set (run under gpurun, one GPU): bench lines, launch list, ncu --set full summaries.
# Every ncu pass runs only after the same command has exited 0 without ncu.
TAG=${1:-r2}
set -x
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
# launch list of one config-2 step (200 clusters keep the capture short; shares are size-independent)
python bench.py --clusters 200 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1 && \
  bash tools/launch_list.sh ${TAG} --clusters 200 > gpurun_out/ll_${TAG}.txt 2>&1
# full capture of the config-2 kernels
bash tools/ncu_block.sh ${TAG} > gpurun_out/ncu_${TAG}.txt 2>&1
python profiles/ncu_summary.py gpurun_out/raw_${TAG}.csv > gpurun_out/ncu_full_summary_${TAG}.txt 2>&1
# full capture of the sample-sliced kernels on the config-4 shape (10,000 genomes)
python bench.py --config 4 --clusters 48 --batch-clusters 48 --steps 1 --no-e2e --no-extra --no-cpu-baseline > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on \
  -k regex:"kA_block_aggregate|kB1_local|kB4_link|kB5_emit|k4_probe|k4_commit" -c 12 \
  -o gpurun_out/prof_${TAG}_cfg4 -f \
  python bench.py --config 4 --clusters 48 --batch-clusters 48 --steps 1 --no-e2e --no-extra --no-cpu-baseline > gpurun_out/ncu_${TAG}_cfg4.log 2>&1
ncu -i gpurun_out/prof_${TAG}_cfg4.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_cfg4.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/raw_${TAG}_cfg4.csv > gpurun_out/ncu_full_summary_${TAG}_cfg4.txt 2>&1
ls -la gpurun_out/*${TAG}*
