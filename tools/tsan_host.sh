#!/bin/sh
# The threaded host entry points under ThreadSanitizer, driven from two small C++ programs (no
# Python, no GPU): genome loading, look-ups and packing of a cut, the ASCII cut, the table reader
# (tools/tsan/feeder_threads.cpp); the kmers_to_hashes formatter on both row layouts, the pattern
# formatter with and without NaN cells, the id text, the compact positional formatter
# (tools/tsan/format_threads.cpp).  usage: sh tools/tsan_host.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${TMPDIR:-/tmp}/pf_tsan
mkdir -p "$OUT/gffs"
python - "$OUT" <<'PY'
import os, sys
import numpy as np
out = sys.argv[1]
rng = np.random.default_rng(1)
acgt = np.frombuffer(b"ACGT", np.uint8)
N, L, G = 300, 1000, 32
for g in range(G):
    name = f"s{g:04d}"
    rows = [f"{name}_c1\tsynth\tCDS\t{1 + c * 1150 + 150}\t{c * 1150 + 150 + L}\t.\t{'+-'[c % 2]}\t0\tID={name}_{c:05d};x=1" for c in range(N)]
    text = acgt[rng.integers(0, 4, N * 1150, dtype=np.uint8)].tobytes().decode()
    with open(os.path.join(out, "gffs", name + ".gff"), "w") as fh:
        fh.write("##gff-version 3\n" + "\n".join(rows) + "\n##FASTA\n>" + name + "_c1\n")
        fh.write("\n".join(text[i:i + 60] for i in range(0, len(text), 60)) + "\n")
with open(os.path.join(out, "table.csv"), "w") as fh:
    fh.write("Gene,Non-unique Gene name,Annotation," + ",".join(f"s{g:04d}" for g in range(400)) + "\n")
    for c in range(400):
        fh.write(f"cl{c},,\"a, b\"," + ",".join(f"s{g:04d}_{c:05d}" if (g + c) % 7 else "" for g in range(400)) + "\n")
PY
cd "$ROOT/tools/tsan"
g++ -O1 -g -std=c++17 -fsanitize=thread -x c++ ../../panfeed_b200/csrc/pf_feeder.cu -x c++ feeder_threads.cpp -o "$OUT/feeder_threads" -lpthread
g++ -O1 -g -std=c++17 -fsanitize=thread -x c++ ../../panfeed_b200/csrc/pf_format.cu -x c++ format_threads.cpp -o "$OUT/format_threads" -lpthread -lz
"$OUT/feeder_threads" "$OUT/gffs" "$OUT/table.csv" > "$OUT/run.log" 2>&1
"$OUT/format_threads" >> "$OUT/run.log" 2>&1
grep -v "^$" "$OUT/run.log" | tail -12
echo "ThreadSanitizer reports: $(grep -c 'ThreadSanitizer' "$OUT/run.log" || true)"
