#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
timeout 600 python scripts/kernel_zoo.py > gpurun_out/kernel_zoo.log 2>&1
timeout 400 python bench.py --workload pretrain --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pretrain.json 2> gpurun_out/bench_pretrain.err
timeout 300 python scripts/pretrain_micro.py > gpurun_out/pm_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pretrain.csv \
   python scripts/pretrain_micro.py > gpurun_out/pm_ncu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/kernel_zoo.log; cut -c1-300 gpurun_out/bench_pretrain.json; cat gpurun_out/pm_plain.log
du -sh gpurun_out
