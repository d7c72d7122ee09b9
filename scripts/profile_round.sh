#!/bin/bash
# evidence refresh: tests, graph-timed kernel zoo, bench lines of all workloads, launch lists, ncu tables
mkdir -p gpurun_out /tmp/ncu
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python scripts/kernel_zoo.py > gpurun_out/kernel_zoo.log 2>&1; grep -v "bn=" gpurun_out/kernel_zoo.log | grep -E "row_stats|folded|CLIP|layernorm_fwd CLIP"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_linear.json 2> gpurun_out/bench_linear.err
for w in qformer xattn; do timeout 400 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; done
timeout 400 python bench.py --workload pretrain --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pretrain.json 2> gpurun_out/bench_pretrain.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cut -c1-200 gpurun_out/bench_*.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_linear.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu.log 2>&1
timeout 300 python scripts/pretrain_micro.py > gpurun_out/pm_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pretrain.csv \
   python scripts/pretrain_micro.py > gpurun_out/pm_ncu.log 2>&1
for sel in layernorm row_stats "attention" "gemm CLIP" "gemm GPT-2 c_fc"; do
  tag=$(echo "$sel" | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:vlk:: -c 40 \
     -o /tmp/ncu/$tag -f python scripts/kernel_zoo.py --once --only "$sel" > gpurun_out/ncu_$tag.log 2>&1
  ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > /tmp/ncu/$tag.csv 2>/dev/null
  python scripts/ncu_compact.py /tmp/ncu/$tag.csv > gpurun_out/ncu_$tag.csv 2>> gpurun_out/ncu_$tag.log
done
du -sh gpurun_out
