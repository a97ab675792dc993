"""Per-stage wall clock of pf_execute on a few batches of a BASELINE shape (PF_DEBUG_TIME=1).
usage: PF_DEBUG_TIME=1 python tools/stage_times.py SAMPLES CLUSTERS_PER_BATCH [BATCHES] [cm]"""
import sys, time, json
sys.path.insert(0, '.')
from panfeed_b200 import capi
S, C = int(sys.argv[1]), int(sys.argv[2])
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cm = len(sys.argv) > 4 and sys.argv[4] == "cm"
ctx = capi.Context(31, S, consider_missing=cm, maf=0.01)
arena = capi.SynthArena()
for b in range(nb):
    hb = capi.synth_batch(0, 20261023, S, C, first_cluster=b * C, total_clusters=nb * C, gene_len=1200, arena=arena)
    t0 = time.perf_counter(); ctx.upload(hb); t1 = time.perf_counter()
    print(f"--- batch {b}: upload {1e3*(t1-t0):.1f} ms", file=sys.stderr)
    ctx.execute(); st = ctx.stats(); t2 = time.perf_counter()
    r = ctx.collect(copy=False); t3 = time.perf_counter()
    print(json.dumps({"batch": b, "execute_wall_ms": round(1e3*(t2-t1), 2), "collect_ms": round(1e3*(t3-t2), 2),
                      **{k: round(st[k], 3) for k in ("ms_extract", "ms_sort", "ms_count", "ms_dedup", "ms_total")},
                      "rows": int(len(r["row_cluster"])), "patterns": st["kmer_patterns"]}))
ctx.close()
