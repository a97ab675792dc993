#!/bin/sh
# The host-only translation units (feeder, table reader, formatters, packer, TSV filter) rebuilt
# with AddressSanitizer + UBSan and linked with the regular pf_api.o, then the CPU tests of the
# native host code against that library (PF_LIB_PATH).  No GPU needed.  usage: sh tools/asan_host.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${TMPDIR:-/tmp}/pf_asan
mkdir -p "$OUT"
make -s -C "$ROOT/panfeed_b200/csrc"
for f in pf_format pf_feeder pf_tsv; do
  g++ -O1 -g -std=c++17 -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -x c++ \
      -c "$ROOT/panfeed_b200/csrc/$f.cu" -o "$OUT/$f.o"
done
g++ -shared -fsanitize=address,undefined -o "$OUT/libpanfeed_b200.so" "$OUT"/pf_format.o "$OUT"/pf_feeder.o \
    "$OUT"/pf_tsv.o "$ROOT/panfeed_b200/csrc/pf_api.o" -L/usr/local/cuda/lib64 -lcudart_static -lpthread -lz -ldl -lrt
cd "$ROOT"
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 \
PF_LIB_PATH="$OUT/libpanfeed_b200.so" python -m pytest tests/test_feeder_native.py tests/test_format_native.py \
    tests/test_pack_native.py tests/test_cli_host.py tests/test_postgwas.py tests/test_capi_symbols.py \
    tests/test_sharded_host.py -q -p no:cacheprovider -s > "$OUT/run.log" 2>&1 || { tail -40 "$OUT/run.log"; exit 1; }
tail -1 "$OUT/run.log"
echo "sanitizer reports: $(grep -c -E 'runtime error|AddressSanitizer' "$OUT/run.log" || true)"
