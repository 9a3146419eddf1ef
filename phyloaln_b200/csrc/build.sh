#!/bin/bash
# Builds phyloaln_b200/libphyloaln_b200.so for sm_100a (B200). No FMA contraction anywhere:
# the float stages must reproduce the canonical evaluation order bit for bit.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libphyloaln_b200.so
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-O2 -shared \
  -o $OUT api.cu profile.cpp ingest.cpp -lz "$@"
echo "built $OUT"
