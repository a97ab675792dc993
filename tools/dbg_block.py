import sys, time
sys.path.insert(0, ".")
from panfeed_b200 import capi
S, C, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
hb = capi.synth_batch(0, 20261018, S, C, total_clusters=C, gene_len=L)
ctx = capi.Context(31, S, maf=0.01)
t = time.time()
ctx.submit(hb); r = ctx.collect(); st = ctx.stats()
print("engine", st["engine"], "rows", st["rows"], "uniq", st["unique_kmers"], "partials", st["partial_rows"], "slots", st["block_slots"], "%.2fs" % (time.time() - t))
ctx.close()
