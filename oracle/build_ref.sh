#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference into oracle/_ref/ (git-ignored).
#
# What it does (only when /root/reference is present, i.e. in the authoring container):
#   1. copies submodules/diff-gaussian-rasterization to a scratch dir under /tmp (the reference
#      tree is read-only) and builds its torch extension for sm_100a with the reference's own
#      setup.py; the only portability fix is forcing `#include <cstdint>` from the command line
#      (rasterizer_impl.h uses uint32_t/uintptr_t without it under gcc 13) — no source edits;
#   2. installs the resulting package (`diff_gaussian_rasterization/__init__.py` + `_C*.so`)
#      into oracle/_ref/ext/;
#   3. installs the reference's pure-Python pipeline (utils/, gaussian_splatting/, main.py,
#      configs/) into oracle/_ref/pipeline/ exactly as `pip install --target` would, so that
#      `bench.py --impl reference` and the parity tests can drive the reference's own code path
#      on the GPU box (where /root/reference does not exist).
# Nothing under oracle/_ref/ is tracked by git; nothing in the product imports it.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GSEVT_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/submodules/diff-gaussian-rasterization" ]; then
  echo "[build_ref] $REF not present; keeping prebuilt oracle/_ref as is"; exit 0
fi
if [ -f "$OUT/ext/diff_gaussian_rasterization/__init__.py" ] && ls "$OUT"/ext/diff_gaussian_rasterization/_C*.so >/dev/null 2>&1 \
   && [ -f "$OUT/pipeline/utils/tracker.py" ] && [ "${1:-}" != "--force" ]; then
  echo "[build_ref] oracle/_ref already built"; exit 0
fi
SCR="$(mktemp -d /tmp/gsevt_ref_build.XXXXXX)"
cp -r "$REF/submodules/diff-gaussian-rasterization" "$SCR/dgr"
chmod -R u+w "$SCR/dgr"
( cd "$SCR/dgr" && NVCC_APPEND_FLAGS="-include cstdint -lineinfo" TORCH_CUDA_ARCH_LIST=10.0a MAX_JOBS=8 \
    python setup.py build_ext --inplace > "$SCR/build.log" 2>&1 ) || { tail -50 "$SCR/build.log"; exit 1; }
mkdir -p "$OUT/ext/diff_gaussian_rasterization" "$OUT/pipeline" "$OUT/obj"
cp "$SCR/dgr/diff_gaussian_rasterization/__init__.py" "$OUT/ext/diff_gaussian_rasterization/"
cp "$SCR"/dgr/diff_gaussian_rasterization/_C*.so "$OUT/ext/diff_gaussian_rasterization/"
# keep the object files for SASS inspection (cuobjdump -sass oracle/_ref/obj/forward.o)
find "$SCR/dgr/build" -name '*.o' -path '*cuda_rasterizer*' -exec cp {} "$OUT/obj/" \;
for d in utils gaussian_splatting configs; do rm -rf "$OUT/pipeline/$d"; cp -r "$REF/$d" "$OUT/pipeline/$d"; done
cp "$REF/main.py" "$OUT/pipeline/main.py"
chmod -R u+w "$OUT/pipeline"
rm -rf "$SCR"
echo "[build_ref] done -> $OUT"
