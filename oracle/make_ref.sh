#!/bin/sh
# ORACLE — TEST INFRASTRUCTURE ONLY.
# "Installs" the UNMODIFIED reference package into oracle/_ref/ so that the reference arm of
# bench.py (`--impl reference`) and the tests can import the reference's own hot functions
# (panfeed/panfeed.py:23-235) on the GPU box, where /root/reference does not exist.  The
# reference is a pure-Python package: installing it is copying its package directory (its
# hatchling build backend is not in this image, so `pip install --target` cannot run).
# oracle/_ref/ is git-ignored (never committed) but travels with gpurun snapshots, like a
# built .so.  The third-party `pyfaidx` the reference imports at module top is not in the
# image: the test-only stand-in of tests/golden/pyfaidx_standin is placed next to it.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${PANFEED_REFERENCE:-/root/reference}"
if [ ! -d "$REF/panfeed" ]; then
  echo "make_ref: $REF/panfeed not found; keeping whatever is in $HERE/_ref" >&2
  exit 0
fi
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref"
cp -r "$REF/panfeed" "$HERE/_ref/panfeed"
cp "$HERE/../tests/golden/pyfaidx_standin/pyfaidx.py" "$HERE/_ref/pyfaidx.py"
find "$HERE/_ref" -name __pycache__ -type d -exec rm -rf {} + 2>/dev/null || true
( cd "$REF" && find panfeed -name '*.py' | sort | xargs sha256sum ) > "$HERE/_ref/SHA256SUMS"
echo "make_ref: reference package $(sed -n "s/__version__ = //p" "$REF/panfeed/__init__.py") -> $HERE/_ref"
