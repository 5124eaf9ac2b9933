#!/bin/bash
# Builds the reference's own CUDA extension (gq_cuda_extension: the legacy `extension_cpp::gq` score kernel)
# from its sources WHERE THEY LIE under /root/reference into oracle/_ref/gq_cuda_ref.so -- two files, compiled with
# nvcc directly against the torch headers of this interpreter (the reference's setup.py is not run).  The result is a
# plain shared object whose static initialisers register the operator when it is loaded with
# torch.ops.load_library(); it is TEST INFRASTRUCTURE: only tests/ and tests/diag/ load it, in a process that has
# not imported gq_cuda (both register `extension_cpp::gq`).  oracle/_ref/ is git-ignored; the built file travels to
# the GPU box with the tree.  No reference source is copied.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=/root/reference/gq_cuda_extension/gq_cuda/csrc
OUT="$HERE/_ref/gq_cuda_ref.so"
[ -d "$REF" ] || { echo "build_ref.sh: $REF not present (GPU box): keeping the prebuilt $OUT"; exit 0; }
if [ -f "$OUT" ] && [ "$OUT" -nt "$REF/cuda/gq_cuda.cu" ] && [ "$OUT" -nt "$REF/gq_cuda.cpp" ]; then exit 0; fi
mkdir -p "$HERE/_ref"
PY=${PYTHON:-python}
read -r TORCH_INC TORCH_LIB PY_INC ABI <<<"$($PY - <<'PYEOF'
import sysconfig, torch
from torch.utils import cpp_extension as ce
inc = ce.include_paths()
print(":".join(inc), ce.library_paths()[0], sysconfig.get_paths()["include"], int(torch._C._GLIBCXX_USE_CXX11_ABI))
PYEOF
)"
INCS=""; IFS=':' read -ra P <<<"$TORCH_INC"; for p in "${P[@]}"; do INCS="$INCS -I$p"; done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared \
  -D_GLIBCXX_USE_CXX11_ABI=$ABI -DTORCH_EXTENSION_NAME=_C -DPy_LIMITED_API=0x03090000 \
  $INCS -I"$PY_INC" -L"$TORCH_LIB" -Xlinker -rpath -Xlinker "$TORCH_LIB" \
  -o "$OUT" "$REF/gq_cuda.cpp" "$REF/cuda/gq_cuda.cu" -lc10 -ltorch_cpu -ltorch -lc10_cuda -ltorch_cuda
echo "built $OUT"
