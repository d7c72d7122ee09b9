#!/bin/bash
mkdir -p gpurun_out
for o in 0 1 2; do
  echo "== VLK_ATTN_V4_ORDER=$o"
  VLK_ATTN_V4_ORDER=$o python scripts/kernel_zoo.py --only "attention fwd CLIP" 2>&1 | tail -1
  VLK_ATTN_V4_ORDER=$o python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -m gpu -x -q -k "attention or clip" 2>&1 | tail -2
  VLK_ATTN_V4_ORDER=$o python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-140
done
