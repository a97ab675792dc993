"""Wall-clock of pf_submit + pf_collect on config 2 with the pipelined submit (PF_DEBUG_PIPE=1
prints the host-side split per submit)."""
import sys, time, json
sys.path.insert(0, '.')
from panfeed_b200 import capi
S, C = 500, int(sys.argv[1]) if len(sys.argv) > 1 else 4000
hb = capi.synth_batch(0, 20261020, S, C, total_clusters=C, gene_len=1200, pinned=True)
ctx = capi.Context(31, S, maf=0.01)
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 5):
    ctx.reset_patterns()
    t0 = time.perf_counter(); ctx.submit(hb); t1 = time.perf_counter()
    r = ctx.collect(copy=False); t2 = time.perf_counter()
    st = ctx.stats()
    print(json.dumps({"submit_ms": round((t1-t0)*1e3,1), "collect_ms": round((t2-t1)*1e3,1), "subs": st["sub_batches"],
                      "dev_ms_total": round(st["ms_total"],1), "h2d_ms_sum": round(st["ms_h2d"],1), "d2h_ms": round(st["ms_d2h"],1),
                      "kA": round(st["ms_sort"],1), "kB": round(st["ms_count"],1), "k4": round(st["ms_dedup"],1), "rows": st["rows"]}))
ctx.close()
