"""Host-side wall-clock split of one end-to-end step (pf_upload / pf_execute / pf_collect)."""
import sys, time, json
sys.path.insert(0, '.')
from panfeed_b200 import capi
import torch
S, C = 500, int(sys.argv[1]) if len(sys.argv) > 1 else 4000
hb = capi.synth_batch(0, 20261020, S, C, total_clusters=C, gene_len=1200, pinned=True)
ctx = capi.Context(31, S, maf=0.01)
for rep in range(4):
    ctx.reset_patterns()
    t0 = time.perf_counter(); ctx.upload(hb); t1 = time.perf_counter()
    ctx.execute(); ctx.stats(); t2 = time.perf_counter()
    r = ctx.collect(copy=False); t3 = time.perf_counter()
    st = ctx.stats()
    print(json.dumps({"upload_ms": round((t1-t0)*1e3,1), "execute_ms": round((t2-t1)*1e3,1), "collect_ms": round((t3-t2)*1e3,1),
                      "h2d_ms": round(st["ms_h2d"],1), "d2h_ms": round(st["ms_d2h"],1), "d2h_bytes": r["d2h_bytes"],
                      "h2d_bytes": hb.packed.nbytes + len(hb.seqs)*64}))
ctx.close()
