#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Compiles the reference's own portable-C implementation of the mars hot path,
# from the sources where they lie under $REF_DIR (default /root/reference), into
# oracle/_ref/libmars_ref.so, and stages the shipped .mars model blobs into
# oracle/_ref/models/ so they travel to the GPU box (oracle/_ref/ is git-ignored).
# No reference source is copied into the repository: the single patched file
# (arena-size literal, src/mars/mars_runtime.c:209) lives in a mktemp dir that is
# deleted at exit.  Flags = the reference's Makefile:21 (-O3 -fPIC -funroll-loops)
# plus -ffp-contract=off to pin "no FMA contraction".
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF_DIR="${REF_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF_DIR/src/mars" ]; then
    echo "build_ref: $REF_DIR not present; keeping prebuilt $OUT" >&2
    [ -f "$OUT/libmars_ref.so" ] && exit 0
    exit 3
fi
mkdir -p "$OUT/models"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT

# the one scripted patch: arena size becomes a run-time value (default 8 MiB)
sed 's/model->ddr_size = 8 \* 1024 \* 1024;/model->ddr_size = oracle_ref_ddr_size();/' \
    "$REF_DIR/src/mars/mars_runtime.c" > "$TMP/mars_runtime_patched.c"
if [ "$(diff "$REF_DIR/src/mars/mars_runtime.c" "$TMP/mars_runtime_patched.c" | grep -c '^[<>]')" != "2" ]; then
    echo "build_ref: arena patch did not change exactly one line" >&2; exit 1
fi
diff "$REF_DIR/src/mars/mars_runtime.c" "$TMP/mars_runtime_patched.c" > "$OUT/arena_patch.diff" || true

CFLAGS="-O3 -fPIC -funroll-loops -ffp-contract=off -w -U_FORTIFY_SOURCE -D_FORTIFY_SOURCE=0"
QUIET="-Dprintf=oracle_ref_printf -Dfprintf=oracle_ref_fprintf"
INC="-I$REF_DIR/include -I$REF_DIR/src -I$REF_DIR"
gcc $CFLAGS $QUIET $INC \
    -c "$TMP/mars_runtime_patched.c" -o "$TMP/mars_runtime.o" \
    -include "$HERE/ref_decls.h"
gcc $CFLAGS $QUIET $INC -c "$REF_DIR/src/mars/mxu_conv.c" -o "$TMP/mxu_conv.o"
gcc $CFLAGS $QUIET $INC -c "$REF_DIR/src/mars/mxu_ops.c" -o "$TMP/mxu_ops.o"
gcc $CFLAGS $QUIET $INC -c "$REF_DIR/src/mars/mars_math.c" -o "$TMP/mars_math.o"
gcc $CFLAGS $INC -c "$HERE/ref_shim.c" -o "$TMP/ref_shim.o"
gcc $CFLAGS $QUIET $INC -DSTB_IMAGE_STATIC -DSTB_IMAGE_RESIZE_STATIC \
    -include "$HERE/ref_decls.h" -c "$HERE/ref_post_shim.c" -o "$TMP/ref_post_shim.o"
g++ $CFLAGS $QUIET $INC -DSTB_IMAGE_STATIC -DSTB_IMAGE_RESIZE_STATIC \
    -include "$HERE/ref_decls.h" -c "$HERE/ref_detect_shim.cpp" -o "$TMP/ref_detect_shim.o"
g++ -shared -Wl,-Bsymbolic -o "$OUT/libmars_ref.so" "$TMP"/*.o -lm
cp -f "$REF_DIR"/models/*.mars "$OUT/models/"
echo "build_ref: wrote $OUT/libmars_ref.so and $(ls "$OUT/models" | wc -l) model blobs"

# --- drop-in proof: the reference's own CALLER programs, unmodified ---------------------
# *.b200 : compiled against THIS repository's include/ and linked with libmars_b200.so only
#          (the link line INTEGRATION.md section 2 gives a maintainer);
# *.ref  : the same sources against the reference's headers and its own runtime (libmars_ref.so).
# tests/test_gpu_dropin.py runs both on the GPU box and compares what they print.
REPO="$(cd "$HERE/.." && pwd)"
B200_LIB="$REPO/thingino-accel_b200/lib"
mkdir -p "$OUT/bin"
if [ -f "$B200_LIB/libmars_b200.so" ]; then
    for app in mars_test mars_yolo_test; do
        gcc -O2 -w -I"$REPO/include" -I"$REF_DIR/include" "$REF_DIR/src/mars/$app.c" -o "$OUT/bin/$app.b200" \
            -L"$B200_LIB" -lmars_b200 -Wl,-rpath,'$ORIGIN/../../../thingino-accel_b200/lib' -lm
        gcc -O2 -w $INC "$REF_DIR/src/mars/$app.c" -o "$OUT/bin/$app.ref" \
            -L"$OUT" -lmars_ref -Wl,-rpath,'$ORIGIN/..' -lm
    done
    echo "build_ref: wrote drop-in caller binaries to $OUT/bin"
else
    echo "build_ref: libmars_b200.so not built yet; skipping the drop-in caller binaries" >&2
fi
