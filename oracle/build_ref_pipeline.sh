#!/bin/bash
# Builds the reference's own `panSVR` and `deBGA` binaries into oracle/_ref/ (SURVEY.md section 8c
# recipe).  TEST INFRASTRUCTURE ONLY: they are the SAM-level oracle and the CPU baseline.
# The reference tree is read-only and is never copied into the repo: a scratch copy under
# $TMPDIR is compiled and only the two executables are kept.
set -euo pipefail
REF=${REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/src" ] || { echo "reference sources not present; keeping prebuilt oracle/_ref (if any)"; exit 0; }
if [ -x "$OUT/panSVR" ] && [ -x "$OUT/deBGA" ] && [ "${1:-}" != "--force" ]; then echo "oracle/_ref/panSVR and deBGA already built"; exit 0; fi
mkdir -p "$OUT"
X="$(mktemp -d "${TMPDIR:-/tmp}/pansvr_ref.XXXXXX")"
trap 'rm -rf "$X"' EXIT
cp -r "$REF/." "$X/"
chmod -R u+w "$X"
# vendored htslib includes <lzma.h> unconditionally; its own stub satisfies it (no lzma symbol is linked)
mkdir -p "$X/shim" && echo '#include "../src/htslib/os/lzma_stub.h"' > "$X/shim/lzma.h"
# the default make goal of Release/ resolves to a `clean` rule: name `all`
( cd "$X/Release" && CPATH="$X/shim" make -j"$(nproc)" all >"$X/build_pansvr.log" 2>&1 ) || { tail -30 "$X/build_pansvr.log"; exit 1; }
cp "$X/Release/panSVR" "$OUT/panSVR"
# gcc >= 10 defaults to -fno-common, deBGA has duplicate tentative definitions
( cd "$X/deBGA_release/src" && make CC="gcc -fcommon" >"$X/build_debga.log" 2>&1 ) || { tail -30 "$X/build_debga.log"; exit 1; }
DEBGA="$(find "$X/deBGA_release" -maxdepth 2 -type f -name deBGA -perm -u+x | head -1)"
cp "$DEBGA" "$OUT/deBGA"
echo "built oracle/_ref/panSVR and oracle/_ref/deBGA"
