#!/bin/bash
# Diagnostic build with per-CTA cycle counters (LFD_TILE_TIMING) -> lentil_b200/liblentil_b200_tt.so.  Development aid.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="$ROOT/lentil_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DLFD_NO_FAST_MATH \
     -DLFD_TILE_TIMING "$@" -o "$ROOT/lentil_b200/liblentil_b200_tt.so" \
     "$SRC"/capi.cu "$SRC"/mft_c128.cu "$SRC"/mft_folded.cu "$SRC"/mft_czt.cu "$SRC"/mft_c64.cu "$SRC"/pupil_prep.cu "$SRC"/accum.cu \
     "$SRC"/fit_tilt.cu "$SRC"/detector_ops.cu "$SRC"/rescale.cu
