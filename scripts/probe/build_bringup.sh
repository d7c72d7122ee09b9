#!/bin/bash
# Bring-up build of the attention kernels (-DVLK_BRINGUP: phase timestamps) -> scripts/probe/libvlk_attn_bringup.so
set -e
cd "$(dirname "$0")/../../gpt2-vision-language_b200/csrc"
mkdir -p /tmp/build_bringup
for f in api attention_api attention_flash attention_pair attention_simt attention_tcgen05; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr -DVLK_BRINGUP -c $f.cu -o /tmp/build_bringup/$f.o 2>/dev/null &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scripts/probe/libvlk_attn_bringup.so /tmp/build_bringup/*.o -lcudart
