#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_linear.json 2> gpurun_out/bench_linear.err; cut -c1-200 gpurun_out/bench_linear.json
