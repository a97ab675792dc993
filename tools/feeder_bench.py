"""Feeding rate (host only): Python feeder (input.py + cluster_cutter) against the native feeder
(pf_feeder_*), both up to and including the packed batch, on a synthetic pangenome written to a
temporary directory.  usage: python tools/feeder_bench.py [genomes] [clusters] [gene_len]"""
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd

sys.path.insert(0, ".")
from panfeed_b200 import feeder as nf          # noqa: E402
from panfeed_b200 import input as pyin         # noqa: E402
from panfeed_b200 import packer, panfeed       # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 200
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
L = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
rng = np.random.default_rng(1)
tmp = tempfile.mkdtemp(prefix="pf_feeder_bench_")
gffdir = os.path.join(tmp, "gffs")
os.mkdir(gffdir)
anc = rng.choice(list("ACGT"), (N, L))
cells = {}
for g in range(G):
    name = f"s{g:04d}"
    rows, seq = [], []
    pos = 1
    col = []
    for c in range(N):
        if rng.random() < 0.1:
            col.append(None)
            continue
        q = anc[c].copy()
        m = rng.random(L) < 0.01
        q[m] = rng.choice(list("ACGT"), int(m.sum()))
        spacer = "".join(rng.choice(list("ACGT"), 150))
        strand = "+" if rng.random() < 0.5 else "-"
        rows.append(f"{name}_c1\tsynth\tCDS\t{pos + 150}\t{pos + 149 + L}\t.\t{strand}\t0\tID={name}_{c:05d};x=1")
        seq.append(spacer + "".join(q))
        pos += 150 + L
        col.append(f"{name}_{c:05d}")
    cells[name] = col
    text = "".join(seq)
    with open(os.path.join(gffdir, name + ".gff"), "w") as fh:
        fh.write("##gff-version 3\n" + "\n".join(rows) + "\n##FASTA\n>" + name + "_c1\n")
        fh.write("\n".join(text[i:i + 60] for i in range(0, len(text), 60)) + "\n")
csv = os.path.join(tmp, "gene_presence_absence.csv")
with open(csv, "w") as fh:
    fh.write("Gene,Non-unique Gene name,Annotation," + ",".join(cells) + "\n")
    for c in range(N):
        fh.write(f"cl{c},,x," + ",".join(cells[g][c] or "" for g in cells) + "\n")
t0 = time.perf_counter()
table = pd.read_csv(csv, sep=",", index_col=0, low_memory=False).drop(columns=["Non-unique Gene name", "Annotation"])
t1 = time.perf_counter()
ntable = nf.PanarooTable(csv)
t2 = time.perf_counter()
print(f"pangenome table ({N} x {G}): pandas read_csv {t1 - t0:.2f} s, library (PanarooTable) {t2 - t1:.3f} s")
stroi = set(list(cells)[:5])
filelist, fastalist = pyin.what_are_my_inputfiles(gffdir, None)
bases = 0

t0 = time.perf_counter()
data = pyin.prep_data_n_fasta(filelist, fastalist, gffdir, None, None)
t1 = time.perf_counter()
pcs = [panfeed.cluster_cutter(x, 31, stroi, False, True, False, None)[1]
       for x in pyin.iter_gene_clusters(table, data, 100, 100, False, True)]
t2 = time.perf_counter()
hb, _, _ = packer.pack_batch(pcs)
t3 = time.perf_counter()
bases = int(hb.seqs["len"].sum())
print(f"python feeder: read {t1 - t0:.2f} s, cut {t2 - t1:.2f} s, pack {t3 - t2:.2f} s  "
      f"-> {bases / (t3 - t1) / 1e6:.1f} Mbases/s cut+pack ({len(hb.seqs)} sequences, {bases} bases)")

t0 = time.perf_counter()
native, index = nf.prep_feeder(filelist, fastalist, gffdir, None)
t1 = time.perf_counter()
npcs = [x[1] for x in nf.iter_packed_clusters(table, native, index, 100, 100, False, stroi, 31, True, False)]
t2 = time.perf_counter()
hb2, _, _ = packer.pack_batch(npcs)
t3 = time.perf_counter()
print(f"native feeder: read {t1 - t0:.2f} s, cut {t2 - t1:.2f} s, pack {t3 - t2:.2f} s  "
      f"-> {bases / (t3 - t1) / 1e6:.1f} Mbases/s cut+pack")
assert (hb.packed == hb2.packed).all() and hb.seqs.tobytes() == hb2.seqs.tobytes()
print("batches identical")

t1 = time.perf_counter()
got = [x[1] for x in nf.iter_packed_batches(ntable, native, index, 100, 100, False, stroi, 31, True, False)]
t2 = time.perf_counter()
nb = sum(int(b.hb.seqs["len"].sum()) for b in got)
assert nb == bases
assert (np.concatenate([b.hb.packed for b in got]) == hb.packed).all()
print(f"native feeder, whole batches from the library's table (iter_packed_batches, {len(got)} batches): cut + pack {t2 - t1:.2f} s "
      f"-> {bases / (t2 - t1) / 1e6:.1f} Mbases/s")
import shutil
shutil.rmtree(tmp)
