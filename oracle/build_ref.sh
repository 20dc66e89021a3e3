#!/usr/bin/env bash
# TEST INFRASTRUCTURE — stages the real reference (andykee/lentil, pure Python) under oracle/_ref/ so that it travels to
# the GPU box with the snapshot (oracle/_ref/ is git-ignored: reference sources never enter this repository's history).
#
# Used only as the checker / CPU baseline:
#   * bench.py --impl reference and bench.py's cpu_baseline leg time it on the host cores (kind: "reference"),
#   * tests/ import it to test lentil_b200.patch.enable() against the real package and to cross-check the oracle port,
#     and run the reference's OWN test files (staged beside it) with its dft2 / idft2 rebound to the CUDA library.
# Nothing under lentil_b200/ imports it.
#
# The copy is pinned: every file's sha256 must match oracle/ref_manifest.sha256 (written once with --pin from the
# reference tree this work was developed against, lentil 0.8.8).  Runs only where /root/reference exists (the build
# container); on the GPU box the already-staged copy is used as is.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${LENTIL_REFERENCE:-/root/reference}"
DST="$HERE/_ref"
MANIFEST="$HERE/ref_manifest.sha256"

if [ ! -d "$SRC/lentil" ]; then
    if [ -d "$DST/lentil" ]; then echo "oracle/_ref: reference tree absent, keeping the staged copy"; exit 0; fi
    echo "oracle/_ref: no reference tree at $SRC and nothing staged" >&2; exit 1
fi
# the package and its own test-suite (tests/test_gpu_reference_own_tests.py runs the latter against the patched package)
list_files() { (cd "$SRC" && find lentil tests -name '*.py' -type f | LC_ALL=C sort); }
if [ "${1:-}" = "--pin" ]; then
    (cd "$SRC" && list_files | xargs sha256sum) > "$MANIFEST"
    echo "pinned $(wc -l < "$MANIFEST") files"
fi
(cd "$SRC" && sha256sum --quiet -c "$MANIFEST")
rm -rf "$DST"
mkdir -p "$DST/lentil"
list_files | while read -r f; do
    mkdir -p "$DST/$(dirname "$f")"
    cp "$SRC/$f" "$DST/$f"
done
(cd "$DST" && sha256sum --quiet -c "$MANIFEST")
echo "oracle/_ref: staged $(wc -l < "$MANIFEST") files of lentil $(grep -o "__version__ = .*" "$DST/lentil/__init__.py" | head -1)"
