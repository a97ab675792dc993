"""Throughput of the library's host-side text formatters on synthetic results (no GPU needed):
kmers_to_hashes rows, hashes_to_patterns rows (with and without NaN cells), kmers.tsv rows from
the compact positional form.  usage: python tools/format_bench.py [threads]   (0 = all cores)"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from panfeed_b200 import capi  # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rng = np.random.default_rng(0)


def best(fn, reps=3):
    out, dt = None, 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        dt = min(dt, time.perf_counter() - t0)
    return out, dt


# kmers_to_hashes: 4 M rows in 1,000 clusters (4,000 rows per cluster, as BASELINE config #2 has)
n, nc = 4_000_000, 1000
r = {"cluster_pattern": rng.integers(0, 100, nc).astype(np.uint32),
     "row_cluster": np.sort(rng.integers(0, nc, n)).astype(np.uint32),
     "row_kmer": rng.integers(0, 1 << 62, n).astype(np.uint64),
     "row_pattern": rng.integers(0, 50000, n).astype(np.uint32)}
kid = capi.base64_ids(rng.integers(0, 256, (50000, 16)).astype(np.uint8))
cid = capi.base64_ids(rng.integers(0, 256, (100, 16)).astype(np.uint8))
tags = [f"group_{i}".encode() for i in range(nc)]
(text, _), dt = best(lambda: capi.format_kmer_rows(r, 31, tags, kid, cid, n_threads=nt, raw=True))
print(f"kmers_to_hashes: {n / dt / 1e6:.1f} M rows/s, {len(text) / dt / 1e6:.0f} MB/s ({len(text) / n:.0f} B/row)")

# hashes_to_patterns: 400,000 patterns of 500 samples
S, n = 500, 400_000
W = (S + 31) // 32
words = rng.integers(0, 1 << 32, (n, W), dtype=np.uint64).astype(np.uint32)
ids = capi.base64_ids(rng.integers(0, 256, (n, 16)).astype(np.uint8))
text, dt = best(lambda: capi.format_patterns(words, S, ids, n_threads=nt, raw=True))
print(f"hashes_to_patterns: {n / dt / 1e6:.2f} M patterns/s, {len(text) / dt / 1e6:.0f} MB/s ({len(text) / n:.0f} B/row)")
pres = rng.integers(0, 1 << 32, (n, W), dtype=np.uint64).astype(np.uint32)
text, dt = best(lambda: capi.format_patterns(words, S, ids, pres, n_threads=nt, raw=True))
print(f"hashes_to_patterns, half of the cells NaN: {n / dt / 1e6:.2f} M patterns/s, {len(text) / dt / 1e6:.0f} MB/s")

# kmers.tsv: 4,000 target sequences of 1,200 bases
nseq, L, k = 4000, 1200, 31
words_per = (L + 63) // 64 * 2
packed = rng.integers(0, 1 << 62, nseq * words_per).astype(np.uint64)
seqs = np.zeros(nseq, capi.SEQ_DTYPE)
seqs["base_off"] = np.arange(nseq, dtype=np.uint64) * (words_per * 32)
seqs["len"], seqs["flags"], seqs["start"], seqs["end"], seqs["offset"] = L, capi.PF_SEQ_TARGET, 1000, 2200, 100
seqs["cluster"], seqs["sample"] = np.arange(nseq) // 400, np.arange(nseq) % 400
seqs["strand"] = np.where(np.arange(nseq) % 2, 1, -1)
hb = capi.HostBatch(packed, seqs, np.zeros(10, capi.CLUSTER_DTYPE), np.zeros((10, 13), np.uint32))
bits = rng.integers(0, 1 << 32, packed.size, dtype=np.uint64).astype(np.uint32)
leads = [f"group_{i // 400}\tstrain{i % 400}\tgene_{i:06d}\tcontig_1\t1\t".encode() for i in range(nseq)]
text, dt = best(lambda: capi.format_positions_compact(hb, bits, k, True, leads, n_threads=nt, raw=True))
rows = nseq * (L - k + 1)
print(f"kmers.tsv: {rows / dt / 1e6:.1f} M rows/s, {len(text) / dt / 1e6:.0f} MB/s ({len(text) / rows:.0f} B/row)")
