import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from panfeed_b200 import capi
for (S, C, cm) in [(10000, 24, False), (50000, 4, True), (2000, 200, False)]:
    t0 = time.time()
    hb = capi.synth_batch(0, 20261018 + 4, S, C, total_clusters=C, gene_len=1200)
    tg = time.time() - t0
    ctx = capi.Context(31, S, consider_missing=cm, maf=0.01)
    ctx.upload(hb)
    for rep in range(3):
        ctx.reset_patterns()
        t0 = time.time(); ctx.execute(); st = ctx.stats(); dt = time.time() - t0
    r = ctx.collect(copy=False)
    st = ctx.stats()
    print(json.dumps({"S": S, "C": C, "cm": cm, "gen_s": round(tg,1), "bases": hb.n_bases, "exec_ms": round(dt*1e3,1),
      "Gbases_s": round(hb.n_bases/dt/1e9,2), "passes": st["sort_passes"], "U": st["unique_kmers"], "rows": st["rows"], "pat": st["kmer_patterns"],
      "ms": {k: round(v,2) for k,v in st.items() if k.startswith("ms_")}}))
    ctx.close()
