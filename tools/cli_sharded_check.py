"""The sharded CLI against the single-process CLI on a pangenome large enough for several GPU
batches: `torchrun --nproc-per-node N -m panfeed_b200 ...` must write the same three files (as
sets of lines) as one process.  usage: python tools/cli_sharded_check.py [ranks] [genomes] [clusters]"""
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 2
G = int(sys.argv[2]) if len(sys.argv) > 2 else 120
N = int(sys.argv[3]) if len(sys.argv) > 3 else 400
L = 900
rng = np.random.default_rng(3)
tmp = tempfile.mkdtemp(prefix="pf_sharded_check_")
gffdir = os.path.join(tmp, "gffs")
os.mkdir(gffdir)
anc = rng.integers(0, 4, (N, L))
founders = [np.where(rng.random((N, L)) < 0.01, (anc + rng.integers(1, 4, (N, L))) & 3, anc) for _ in range(6)]
lut = np.frombuffer(b"ACGT", np.uint8)
names = [f"s{g:04d}" for g in range(G)]
cols = {}
for name in names:
    rows, parts, col, pos = [], [], [], 1
    for c in range(N):
        if rng.random() < (0.05 if c % 3 else 0.5):
            col.append("")
            continue
        q = founders[int(rng.integers(6))][c].copy()
        m = rng.random(L) < 0.002
        q[m] = (q[m] + rng.integers(1, 4, int(m.sum()))) & 3
        strand = "+" if rng.random() < 0.5 else "-"
        rows.append(f"{name}_c1\tsynth\tCDS\t{pos + 120}\t{pos + 119 + L}\t.\t{strand}\t0\tID={name}_{c:05d};x=1")
        parts.append(lut[rng.integers(0, 4, 120)].tobytes().decode() + lut[q].tobytes().decode())
        pos += 120 + L
        col.append(f"{name}_{c:05d}")
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
open(os.path.join(tmp, "targets.txt"), "w").write("\n".join(names[::17]) + "\n")

ok = True
for label, extra in (("first pass", []), ("cluster-absent + targets + gzip", ["--consider-missing", "--compress", "--targets",
                                                                              os.path.join(tmp, "targets.txt")])):
    outs = []
    for mode in ("single", "sharded"):
        out = os.path.join(tmp, f"out_{mode}_{len(extra)}")
        args = ["-g", gffdir, "-p", csv, "-o", out, "--upstream", "60", "--downstream", "60", "--native-feeder"] + extra
        t0 = time.perf_counter()
        if mode == "single":
            cmd = [sys.executable, "-m", "panfeed_b200"] + args
        else:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(R),
                   "--master-addr", "127.0.0.1", "--master-port", "29720", "-m", "panfeed_b200"] + args
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True,
                           env=dict(os.environ, PYTHONPATH=ROOT))
        if r.returncode != 0:
            print(mode, "FAILED", r.stderr[-2000:])
            sys.exit(1)
        outs.append((out, time.perf_counter() - t0))
    import gzip
    for name in ("kmers.tsv", "kmers_to_hashes.tsv", "hashes_to_patterns.tsv"):
        def read(d):
            p = os.path.join(d, name)
            return gzip.open(p + ".gz", "rt").read() if os.path.exists(p + ".gz") else open(p).read()
        a, b = read(outs[0][0]), read(outs[1][0])
        same = a.split("\n")[0] == b.split("\n")[0] and sorted(a.split("\n")) == sorted(b.split("\n"))
        ok &= same
        print(f"{label}: {name}: {len(a.splitlines())} lines, sharded == single: {same}")
    print(f"{label}: single {outs[0][1]:.1f} s, {R} ranks {outs[1][1]:.1f} s (process start-up included)")
shutil.rmtree(tmp)
print("OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
