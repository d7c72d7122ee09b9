#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/kernel_zoo.py --only "gemm" 2>&1 | grep -v "bn="
python scripts/gemm_small_probe.py 2>&1 | grep -E "res=1|16384|16448" | grep -v "bn="
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-140; done
python bench.py --workload pretrain --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-140
