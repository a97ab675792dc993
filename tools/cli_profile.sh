#!/bin/bash
# cProfile of the CLI (native feeder, first pass) on the synthetic pangenome tools/cli_bench.py writes
python - <<'PY'
import cProfile, pstats, sys, os, io, runpy
sys.argv = ["cli_bench", "200", "300", "1000"]
sys.path.insert(0, ".")
# reuse the generator of cli_bench by running it up to the CLI calls, then profile one call
src = open("tools/cli_bench.py").read().split("from panfeed_b200.__main__ import main")[0]
g = {"__name__": "gen"}
exec(compile(src, "cli_bench_gen", "exec"), g)
from panfeed_b200.__main__ import main
tmp, gffdir, csv = g["tmp"], g["gffdir"], g["csv"]
pr = cProfile.Profile()
pr.enable()
main(["-g", gffdir, "-p", csv, "-o", os.path.join(tmp, "out_prof"), "--upstream", "100", "--downstream", "100", "--native-feeder"])
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(35)
print(s.getvalue())
import shutil; shutil.rmtree(tmp)
PY
