#!/bin/sh
# Builds the reference-derived checker objects into oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).
# Only runs where the reference tree is mounted.  The one piece of the pose-graph path the reference vendors in compilable
# form is CSparse (the sparse Cholesky behind g2o's `lm_var` / `gn_var` solvers), inside 3rdtools/g2o-a48ff8c.zip under
# g2o/EXTERNAL/csparse/.  The sources are unpacked to a temporary directory and compiled from there with gcc directly
# (no CMake); nothing from the reference is copied into the repository.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ZIP=/root/reference/3rdtools/g2o-a48ff8c.zip
OUT="$HERE/_ref"
[ -f "$ZIP" ] || { echo "build_ref.sh: $ZIP not found, skipping"; exit 0; }
if [ -f "$OUT/libcsparse_ref.so" ] && [ "$OUT/libcsparse_ref.so" -nt "$ZIP" ]; then exit 0; fi
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
python3 - "$ZIP" "$TMP" <<'PY'
import sys, zipfile
z = zipfile.ZipFile(sys.argv[1])
for n in z.namelist():
    if n.startswith("g2o/EXTERNAL/csparse/") and (n.endswith(".c") or n.endswith(".h")):
        z.extract(n, sys.argv[2])
PY
mkdir -p "$TMP/stub/g2o" "$OUT"
: > "$TMP/stub/g2o/config.h"      # cs_api.h includes g2o/config.h, which CMake would generate; nothing in it is needed
/usr/bin/gcc -O2 -fPIC -shared -I"$TMP/stub" -I"$TMP/g2o/EXTERNAL/csparse" -o "$OUT/libcsparse_ref.so" "$TMP"/g2o/EXTERNAL/csparse/*.c -lm
echo "built $OUT/libcsparse_ref.so"
