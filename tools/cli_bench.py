"""CLI-level timing: `python -m panfeed_b200` end to end (GFF parsing, cutting, packing, GPU,
text output) on a synthetic pangenome written to a temporary directory, first pass and second
pass (--targets = all genomes, --genes = a tenth of the clusters), both feeders.
usage: python tools/cli_bench.py [genomes] [clusters] [gene_len]"""
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 200
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
L = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
rng = np.random.default_rng(1)
tmp = tempfile.mkdtemp(prefix="pf_cli_bench_")
gffdir = os.path.join(tmp, "gffs")
os.mkdir(gffdir)
anc = rng.integers(0, 4, (N, L))
founders = [np.where(rng.random((N, L)) < 0.01, (anc + rng.integers(1, 4, (N, L))) & 3, anc) for _ in range(8)]
lut = np.frombuffer(b"ACGT", np.uint8)
names = [f"s{g:04d}" for g in range(G)]
cols = {}
bases = 0
for name in names:
    rows, parts, col, pos = [], [], [], 1
    for c in range(N):
        if rng.random() < 0.1:
            col.append("")
            continue
        q = founders[int(rng.integers(8))][c].copy()
        m = rng.random(L) < 0.001
        q[m] = (q[m] + rng.integers(1, 4, int(m.sum()))) & 3
        strand = "+" if rng.random() < 0.5 else "-"
        rows.append(f"{name}_c1\tsynth\tCDS\t{pos + 150}\t{pos + 149 + L}\t.\t{strand}\t0\tID={name}_{c:05d};x=1")
        parts.append(lut[rng.integers(0, 4, 150)].tobytes().decode() + lut[q].tobytes().decode())
        pos += 150 + L
        col.append(f"{name}_{c:05d}")
        bases += L + 200
    cols[name] = col
    text = "".join(parts)
    with open(os.path.join(gffdir, name + ".gff"), "w") as fh:
        fh.write("##gff-version 3\n" + "\n".join(rows) + "\n##FASTA\n>" + name + "_c1\n")
        fh.write("\n".join(text[i:i + 60] for i in range(0, len(text), 60)) + "\n")
csv = os.path.join(tmp, "gpa.csv")
with open(csv, "w") as fh:
    fh.write("Gene,Non-unique Gene name,Annotation," + ",".join(names) + "\n")
    for c in range(N):
        fh.write(f"cl{c},,x," + ",".join(cols[n][c] for n in names) + "\n")
open(os.path.join(tmp, "targets.txt"), "w").write("\n".join(names) + "\n")
open(os.path.join(tmp, "genes.txt"), "w").write("\n".join(f"cl{c}" for c in range(0, N, 10)) + "\n")
print(f"pangenome: {G} genomes x {N} clusters, {bases / 1e6:.0f} Mbases to cut (100-bp flanks)")

from panfeed_b200.__main__ import main  # noqa: E402

for label, extra in (("first pass, python feeder", ["--python-feeder"]), ("first pass, native feeder (default)", []),
                     ("second pass (all targets, 1/10 of the clusters), native feeder",
                      ["--targets", os.path.join(tmp, "targets.txt"), "--genes",
                       os.path.join(tmp, "genes.txt")])):
    out = os.path.join(tmp, "out_" + str(abs(hash(label)) % 10000))
    t0 = time.perf_counter()
    main(["-g", gffdir, "-p", csv, "-o", out, "--upstream", "100", "--downstream", "100"] + extra)
    dt = time.perf_counter() - t0
    size = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out) if os.path.isfile(os.path.join(out, f)))
    frac = 0.1 if "second" in label else 1.0
    print(f"{label}: {dt:.2f} s wall, {bases * frac / dt / 1e6:.1f} Mbases/s, {size / 1e6:.0f} MB of output")
shutil.rmtree(tmp)
