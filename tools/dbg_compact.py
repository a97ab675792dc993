import os, sys
import numpy as np
sys.path.insert(0, '.')
from panfeed_b200 import capi
S, C, L, k = 48, 120, 260, 21
hb = capi.synth_batch(0, 77, S, C, total_clusters=C, gene_len=L, all_targets=True)
hb.seqs["flags"][::3] = 0
def run(pipe):
    if pipe: os.environ["PF_PIPELINE_SEQS"] = "1500"
    else: os.environ["PF_PIPELINE_SEQS"] = "0"
    ctx = capi.Context(k, S, canonical=True, emit_positions=2, maf=0.02)
    ctx.submit(hb); r = ctx.collect(); st = ctx.stats(); ctx.close()
    return r, st
a, sa = run(True); b, sb = run(False)
print("subs", sa["sub_batches"], sb["sub_batches"], "n_pos", a["n_pos"], b["n_pos"], len(a["pos_strand_bits"]), len(b["pos_strand_bits"]))
ba, bb = a["pos_strand_bits"], b["pos_strand_bits"]
bad = []
for i, q in enumerate(hb.seqs):
    if not (q["flags"] & 1): continue
    w0 = int(q["base_off"]) >> 5; nw = (int(q["len"]) - k + 1 + 31) // 32
    if not np.array_equal(ba[w0:w0+nw], bb[w0:w0+nw]): bad.append(i)
print("seqs with different bits:", len(bad), bad[:20])
if bad:
    i = bad[0]; q = hb.seqs[i]; w0 = int(q["base_off"]) >> 5
    print(i, q, ba[w0:w0+8], bb[w0:w0+8])
    first = np.searchsorted(hb.seqs["cluster"], np.arange(C + 1))
    print("cluster of bad", [int(hb.seqs["cluster"][j]) for j in bad[:20]])
