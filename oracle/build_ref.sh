#!/usr/bin/env bash
# Builds the UNMODIFIED reference kNN CUDA extension (torch_knnquery @ 947957e,
# /root/reference/torch_knnquery/src/knnquery.cu) for sm_100a into oracle/_ref/.
# Test infrastructure only: oracle/_ref/ is git-ignored, travels to the GPU box
# with gpurun, and is used by tests/ and bench.py as the *checker*, never shipped.
# No reference source is copied: nvcc reads it where it lies.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${SPF_REFERENCE_ROOT:-/root/reference}/torch_knnquery/src/knnquery.cu"
OUT="$HERE/_ref"
[ -f "$SRC" ] || { echo "reference source not present: $SRC (skipping)"; exit 0; }
mkdir -p "$OUT"
PY=python
TORCH_DIR=$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))')
PYINC=$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')
EXT=$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')
TARGET="$OUT/knnquery_cuda$EXT"
if [ -f "$TARGET" ] && [ "$TARGET" -nt "$SRC" ] && [ "$TARGET" -nt "$HERE/ref_compat_shim.h" ]; then
  echo "up to date: $TARGET"; exit 0
fi
nvcc -O2 -std=c++17 -shared -Xcompiler -fPIC \
  -gencode arch=compute_100a,code=sm_100a \
  -include "$HERE/ref_compat_shim.h" \
  -DTORCH_EXTENSION_NAME=knnquery_cuda -DTORCH_API_INCLUDE_EXTENSION_H \
  -D_GLIBCXX_USE_CXX11_ABI=1 \
  -I"$TORCH_DIR/include" -I"$TORCH_DIR/include/torch/csrc/api/include" -I"$PYINC" \
  -Xcudafe --diag_suppress=20012 -w \
  "$SRC" -o "$TARGET" \
  -L"$TORCH_DIR/lib" -lc10 -ltorch -ltorch_cpu -ltorch_python -lc10_cuda -ltorch_cuda \
  -Xlinker -rpath -Xlinker "$TORCH_DIR/lib"
echo "built $TARGET"
