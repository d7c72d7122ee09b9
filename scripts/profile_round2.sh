#!/bin/bash
# Round-2 evidence refresh in one GPU call: tests, graph-timed kernel zoo, the default bench line (all workloads), the
# reference arm, launch lists of the caption step and of a pretraining micro-step, `ncu --set full` tables per kernel class.
mkdir -p gpurun_out/r02 /tmp/ncu
O=gpurun_out/r02
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest_gpu.log; cat $O/pytest_gpu.log
timeout 900 python scripts/kernel_zoo.py > $O/kernel_zoo.log 2>&1; cp gpurun_out/kernel_zoo.json $O/kernel_zoo.json
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; cut -c1-200 $O/bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-extra-workloads > $O/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_linear.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-extra-workloads > $O/ncu.log 2>&1
timeout 300 python scripts/pretrain_micro.py > $O/pm_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_pretrain.csv \
   python scripts/pretrain_micro.py > $O/pm_ncu.log 2>&1
for sel in layernorm "attention" "gemm CLIP" "gemm GPT-2 c_fc"; do
  tag=$(echo "$sel" | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:vlk:: -c 60 \
     -o /tmp/ncu/$tag -f python scripts/kernel_zoo.py --once --only "$sel" > $O/ncu_$tag.log 2>&1
  ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > /tmp/ncu/$tag.csv 2>/dev/null
  python scripts/ncu_compact.py /tmp/ncu/$tag.csv > $O/ncu_$tag.csv 2>> $O/ncu_$tag.log
done
du -sh $O
