#!/bin/bash
# diagnostic: build experiment variants of the library into scripts/_build/ (never shipped)
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/_build
SRC="vq-vae-from-gaussian-vae_b200/csrc"
i=0
for flags in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 $flags -Xcompiler -fPIC -shared \
    -o scripts/_build/libgq_v$i.so $SRC/gq_api.cu $SRC/gq_kernels.cu $SRC/gq_kl.cu $SRC/gq_tc.cu $SRC/gq_tokens.cu &
  i=$((i+1))
done
wait
ls -la scripts/_build/
