#!/usr/bin/env python
"""Golden outputs of the reference's post-GWAS tools (panfeed-get-clusters, panfeed-get-kmers),
produced by running the UNMODIFIED /root/reference/panfeed/get_clusters.py and get_kmers.py on the
committed golden outputs of the `updownstream` mode plus a synthetic pyseer-like associations
table (p-values derived from the md5 of every pattern hash, so the script is deterministic).

Only works in the build container (needs /root/reference).  Writes
tests/golden/expected/postgwas/: associations.tsv and one .txt.gz per case (stdout of the tool).
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
MODE = "updownstream"
OUT = os.path.join(HERE, "expected", "postgwas")

CASES = {
    "clusters_t0.3": ("get_clusters", ["--threshold", "0.3"]),
    "clusters_t0.05_col": ("get_clusters", ["--threshold", "0.05", "--column", "filter-pvalue"]),
    "clusters_all": ("get_clusters", []),
    "kmers_t0.3": ("get_kmers", ["--threshold", "0.3"]),
    "kmers_t0.3_passing": ("get_kmers", ["--threshold", "0.3", "--only-passing"]),
    "kmers_t0.1_iter2": ("get_kmers", ["--threshold", "0.1", "--clusters-per-iteration", "2"]),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="pf_postgwas_")
    for name in ("kmers.tsv", "kmers_to_hashes.tsv", "hashes_to_patterns.tsv"):
        with gzip.open(os.path.join(HERE, "expected", MODE, name + ".gz"), "rt") as fh, \
                open(os.path.join(tmp, name), "w") as out:
            out.write(fh.read())
    assoc = os.path.join(OUT, "associations.tsv")
    with open(os.path.join(tmp, "hashes_to_patterns.tsv")) as fh, open(assoc, "w") as out:
        out.write("variant\taf\tfilter-pvalue\tlrt-pvalue\tbeta\n")
        next(fh)
        for line in fh:
            h = line.split("\t")[0]
            d = hashlib.md5(h.encode()).hexdigest()
            p1 = (int(d[:8], 16) % 1000) / 1000
            p2 = (int(d[8:16], 16) % 1000) / 1000
            out.write(f"{h}\t0.25\t{p2}\t{p1}\t{int(d[16:18], 16) / 64 - 2}\n")
    env = dict(os.environ, PYTHONPATH=REF)
    for case, (tool, extra) in CASES.items():
        args = ["-a", assoc, "-p", os.path.join(tmp, "kmers_to_hashes.tsv")]
        if tool == "get_kmers":
            args += ["-k", os.path.join(tmp, "kmers.tsv")]
        code = (f"import sys; sys.argv = ['panfeed-{tool}'] + {args + extra!r}; "
                f"from panfeed.{tool} import main; main()")
        res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True)
        with gzip.open(os.path.join(OUT, case + ".txt.gz"), "wt", compresslevel=9) as fh:
            fh.write(res.stdout)
        print(case, len(res.stdout.splitlines()), "lines")
    with open(os.path.join(OUT, "cases.json"), "w") as fh:
        json.dump({"mode": MODE, "cases": {k: [v[0], v[1]] for k, v in CASES.items()}}, fh, indent=1)


if __name__ == "__main__":
    main()
