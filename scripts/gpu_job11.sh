#!/bin/bash
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -3
python scripts/kernel_zoo.py --only "pretrain B=16 H=12 T=1024" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --workload pretrain --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-140
