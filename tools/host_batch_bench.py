"""Host side of one first-pass batch without a GPU: `panfeed._run_batch` + the file writes on a
REPLAYED result of BASELINE config #2's shape (per cluster ~3,600 rows and ~800 new patterns of 500
samples), i.e. what the CLI does between `pf_collect` and the next `pf_submit`.
usage: python tools/host_batch_bench.py [clusters] [samples]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
from panfeed_b200 import capi, panfeed  # noqa: E402

NC = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 500
W = (S + 31) // 32
rng = np.random.default_rng(0)
rows_per, pats_per = 3600, 800
n, npat = NC * rows_per, NC * pats_per


class ReplayContext:
    """Answers collect() with a synthetic result of the right shape (no device)."""
    W, Wk, consider_missing = W, W, False

    def __init__(self):
        self.r = {
            "row_cluster": np.repeat(rng.permutation(NC), rows_per).astype(np.uint32),       # one run per cluster
            "row_kmer": rng.integers(0, 1 << 62, n).astype(np.uint64),
            "row_count": rng.integers(5, 495, n).astype(np.uint32),
            "row_pattern": rng.integers(0, npat, n).astype(np.uint32),
            "wide_row_cluster": np.zeros(0, np.uint32), "wide_row_kmer": np.zeros((0, 2), np.uint64),
            "wide_row_count": np.zeros(0, np.uint32), "wide_row_pattern": np.zeros(0, np.uint32),
            "cluster_pattern": np.arange(NC, dtype=np.uint32), "kmer_pattern_base": 0,
            "new_kmer_patterns": rng.integers(0, 1 << 32, (npat, W), dtype=np.uint64).astype(np.uint32),
            "cluster_pattern_base": 0,
            "new_cluster_patterns": rng.integers(0, 1 << 32, (NC, W), dtype=np.uint64).astype(np.uint32),
            "n_pos": 0, "pos_strand_bits": np.zeros(0, np.uint32)}
        self.dig = {True: rng.integers(0, 256, (NC, 16)).astype(np.uint8),
                    False: rng.integers(0, 256, (npat, 16)).astype(np.uint8)}

    def submit(self, hb):
        pass

    def collect(self, copy=True):
        return self.r

    def pattern_ids(self, ns, first, count):
        return capi.base64_ids(self.dig[bool(ns)][first:first + count])


ctx = ReplayContext()
hb = capi.HostBatch(np.zeros(2, np.uint64), np.zeros(0, capi.SEQ_DTYPE), np.zeros(NC, capi.CLUSTER_DTYPE),
                    np.zeros((NC, W), np.uint32))
idxs = [f"group_{c}" for c in range(NC)]
tmp = tempfile.mkdtemp(prefix="pf_host_batch_")
for rep in range(2):
    store = panfeed.PatternStore()
    with open(os.path.join(tmp, "h2p.tsv"), "w") as hp, open(os.path.join(tmp, "k2h.tsv"), "w") as kh:
        t0 = time.perf_counter()
        pos, pats, rows = panfeed._run_batch(store, ctx, hb, lambda i: None, idxs, S, 31, True, False)
        t1 = time.perf_counter()
        for t in pats:
            panfeed.write_text(hp, t)
        panfeed.write_text(kh, rows)
        hp.flush(); kh.flush()
        t2 = time.perf_counter()
    size = sum(os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp))
    bases = NC * 500 * 1200 * S // 500 // 1 if S == 500 else NC * S * 1200
    print(f"{NC} clusters x {S} samples ({n / 1e6:.1f} M rows, {npat / 1e6:.2f} M new patterns, ~{NC * S * 1200 / 1e6:.0f} Mbases "
          f"of input): ids + formatting {t1 - t0:.2f} s, writing {size / 1e6:.0f} MB {t2 - t1:.2f} s "
          f"-> {NC * S * 1200 / (t2 - t0) / 1e6:.0f} Mbases/s of input")
import shutil
shutil.rmtree(tmp)
