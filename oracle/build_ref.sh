#!/usr/bin/env bash
# Build the UNMODIFIED reference (lmdu/pyfastx v2.3.1) C sources, where they lie under
# /root/reference/src, into oracle/_ref/pyfastx.<abi>.so.  TEST INFRASTRUCTURE ONLY:
# used to pin oracle/fxo.c, to generate tests/golden/ fixtures and as the CPU
# "reference" arm of bench.py.  Nothing is copied into the repo; outputs are git-ignored
# but travel to the GPU box with the gpurun snapshot.
#   zlib    -> system zlib 1.3 (reference pins 1.3.1; same API for the calls used)
#   sqlite3 -> system libsqlite3.so.0 via the prototype shim oracle/ref_shim/sqlite3.h
#   zran    -> naive stand-in oracle/ref_shim/zran.{h,c} (indexed_gzip is not vendored)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${FXO_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF/src not present (GPU box uses the prebuilt oracle/_ref)"; exit 0
fi
mkdir -p "$OUT"
PYINC="$(python3 -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
SUFFIX="$(python3 -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
[ -f "$PYINC/Python.h" ] || PYINC=/usr/include/python3.12
gcc -O2 -shared -fPIC -w -D_FILE_OFFSET_BITS=64 -D_LFS64_LARGEFILE -D_LARGEFILE64_SOURCE \
    -I"$HERE/ref_shim" -I"$PYINC" -I"$REF/src" \
    "$REF"/src/*.c "$HERE/ref_shim/zran.c" \
    -o "$OUT/pyfastx$SUFFIX" -lz -l:libsqlite3.so.0
echo "build_ref: built $OUT/pyfastx$SUFFIX"
