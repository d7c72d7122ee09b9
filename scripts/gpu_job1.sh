#!/bin/bash
# GPU job: parity tests, per-kernel timings, bench of every workload, ncu captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
timeout 600 python scripts/kernel_zoo.py > gpurun_out/kernel_zoo.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_linear.json 2> gpurun_out/bench_linear.err
for w in qformer xattn; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 400 python bench.py --workload pretrain --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pretrain.json 2> gpurun_out/bench_pretrain.err
timeout 300 python scripts/pretrain_micro.py > gpurun_out/pm_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pretrain.csv \
   python scripts/pretrain_micro.py > gpurun_out/pm_ncu.log 2>&1
timeout 300 python scripts/kernel_zoo.py --once > gpurun_out/zoo_once_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:vlk:: \
   -o gpurun_out/kernel_zoo -f python scripts/kernel_zoo.py --once > gpurun_out/zoo_ncu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/kernel_zoo.log; cat gpurun_out/bench_*.json | cut -c1-400
