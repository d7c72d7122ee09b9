#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload pretrain --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-140
python bench.py --workload qformer --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-140
