#!/bin/bash
# Bring-up build of libvlk (-DVLK_BRINGUP: phase timestamps in the attention kernels, wait-time counters and
# work-skipping switches in the GEMM) -> scripts/probe/libvlk_bringup.so.  Never loaded by the package.
# usage: build_bringup.sh [extra nvcc flags, e.g. -DVLK_EXPERIMENT=1] ; output name can be set with OUT=...
set -e
cd "$(dirname "$0")/../../gpt2-vision-language_b200/csrc"
OUT=${OUT:-libvlk_bringup.so}
B=/tmp/build_bringup_${OUT%.so}
mkdir -p $B
for f in *.cu; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr -DVLK_BRINGUP "$@" -c $f -o $B/${f%.cu}.o 2> $B/${f%.cu}.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scripts/probe/$OUT $B/*.o -lcudart
ln -sf $OUT ../../scripts/probe/libvlk_attn_bringup.so 2>/dev/null || true
