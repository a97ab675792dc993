"""Device memory in use after a resident config-2 step and after a pipelined submit."""
import sys
sys.path.insert(0, ".")
import torch
from panfeed_b200 import capi
S, C = 500, 4000
free0, total = torch.cuda.mem_get_info(0)
hb = capi.synth_batch(0, 20261020, S, C, total_clusters=C, gene_len=1200, pinned=True)
ctx = capi.Context(31, S, maf=0.01)
ctx.upload(hb); ctx.execute(); ctx.collect(copy=False)
free1, _ = torch.cuda.mem_get_info(0)
print("resident whole batch: %.1f GB in use" % ((free0 - free1) / 1e9))
ctx.close()
ctx = capi.Context(31, S, maf=0.01)
for _ in range(2):
    ctx.reset_patterns(); ctx.submit(hb); ctx.collect(copy=False)
free2, _ = torch.cuda.mem_get_info(0)
print("pipelined submit (6 sub-batches, 2 slots): %.1f GB in use" % ((free0 - free2) / 1e9))
ctx.close()
