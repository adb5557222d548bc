#!/usr/bin/env bash
# Build the compiled object layer next to libfxg.so (in-tree, git-ignored, travels with gpurun):
#   pyfastx_b200/_fast.<abi>.so    Cython bridge: per-object getters call the C-ABI directly (links libfxg.so)
#   pyfastx_b200/pyfastx.<abi>.so  the object layer (api.py) compiled by Cython into a CPython extension module that
#                                  exports PyInit_pyfastx and registers Fasta / Fastq / Sequence / Read ... like the
#                                  reference's module.c:61-138
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
PKG="$HERE/.."
PY="${PYTHON:-python3}"
INC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
[ -f "$INC/Python.h" ] || INC=/usr/include/python3.12
SUF="$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
# compiled inside a package directory so that Cython gives the modules their qualified names
# (pyfastx_b200.pyfastx: a bare "pyfastx" would shadow an installed reference in sys.modules)
mkdir -p "$TMP/pyfastx_b200"
touch "$TMP/pyfastx_b200/__init__.py"
cp "$HERE/ext/_fast.pyx" "$TMP/pyfastx_b200/_fast.pyx"
{ echo "# cython: language_level=3"; cat "$PKG/api.py"; } > "$TMP/pyfastx_b200/pyfastx.py"
( cd "$TMP" && $PY -m cython -3 pyfastx_b200/_fast.pyx -o _fast.c && $PY -m cython -3 pyfastx_b200/pyfastx.py -o pyfastx.c )
gcc -O2 -fPIC -shared -w -I"$INC" -I"$HERE/../../include" "$TMP/_fast.c" -o "$PKG/_fast$SUF" -L"$PKG" -lfxg -Wl,-rpath,'$ORIGIN'
gcc -O2 -fPIC -shared -w -I"$INC" "$TMP/pyfastx.c" -o "$PKG/pyfastx$SUF"
echo "built $PKG/_fast$SUF and $PKG/pyfastx$SUF"
