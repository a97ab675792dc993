"""Row filter of the post-GWAS joins: pandas in chunks of 100,000 rows (what the reference's
get_clusters.py:90-95 does) against pf_tsv_filter, on a synthetic kmers_to_hashes.tsv.
usage: python tools/postgwas_bench.py [million rows]"""
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd

sys.path.insert(0, ".")
from panfeed_b200 import postgwas      # noqa: E402

M = float(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(M * 1e6)
rng = np.random.default_rng(3)
path = os.path.join(tempfile.mkdtemp(prefix="pf_postgwas_"), "kmers_to_hashes.tsv")
acgt = np.frombuffer(b"ACGT", np.uint8)
with open(path, "wb") as fh:
    fh.write(b"cluster\tk-mer\thashed_pattern\n")
    for a in range(0, n, 500_000):
        m = min(500_000, n - a)
        km = acgt[rng.integers(0, 4, (m, 31))]
        hs = rng.integers(0, 2_000_000, m)
        cl = rng.integers(0, 4000, m)
        rows = [b"group_%d\t%s\tH%020d==" % (cl[i], km[i].tobytes(), hs[i]) for i in range(m)]
        fh.write(b"\n".join(rows) + b"\n")
keys = {"H%020d==" % i for i in rng.choice(2_000_000, 2000, replace=False)}
size = os.path.getsize(path) / 1e6
t0 = time.perf_counter()
chunks = pd.read_csv(path, sep="\t", iterator=True, chunksize=100_000)
h = pd.concat([x[x["hashed_pattern"].isin(keys)] for x in chunks])
t1 = time.perf_counter()
g = postgwas.filter_rows(path, "hashed_pattern", keys)
t2 = time.perf_counter()
assert len(h) == len(g) and (h["k-mer"].values == g["k-mer"].values).all()
print(f"{n} rows, {size:.0f} MB, {len(h)} matches: pandas chunks {t1 - t0:.2f} s ({size / (t1 - t0):.0f} MB/s), "
      f"pf_tsv_filter {t2 - t1:.2f} s ({size / (t2 - t1):.0f} MB/s)")
os.unlink(path)
