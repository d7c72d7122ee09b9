#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-140; done
VLK_CLIP_NO_FUSED_STATS=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-140
