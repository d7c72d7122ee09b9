#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python scripts/kernel_zoo.py > gpurun_out/kernel_zoo.log 2>&1
grep -E "attention|layernorm_bwd|colsum|gemm CLIP out|gemm GPT-2" gpurun_out/kernel_zoo.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_linear.json 2> gpurun_out/bench_linear.err
timeout 400 python bench.py --workload xattn --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_xattn.json 2> gpurun_out/bench_xattn.err
timeout 400 python bench.py --workload qformer --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_qformer.json 2> gpurun_out/bench_qformer.err
timeout 400 python bench.py --workload pretrain --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pretrain.json 2> gpurun_out/bench_pretrain.err
cut -c1-330 gpurun_out/bench_*.json
# ncu --set full of every kernel class at its bench shape; only the raw-page CSV travels back (reports stay on the box)
METRICS='gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|dram__throughput.avg.pct_of_peak_sustained_elapsed|gpu__dram_throughput|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct_of_peak|launch__registers_per_thread|launch__grid_size|launch__block_size|sm__throughput.avg.pct|smsp__cycles_active.avg|sm__cycles_elapsed.avg |sm__cycles_elapsed.max|gpc__cycles_elapsed.max|lts__t_bytes.sum |l1tex__t_bytes.sum '
for sel in layernorm pool33 embed softmax_ce adamw "grad_sumsq" "attention" "gemm CLIP" "gemm GPT-2 c_fc"; do
  tag=$(echo "$sel" | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:vlk:: -c 40 \
     -o /tmp/ncu/$tag -f python scripts/kernel_zoo.py --once --only "$sel" > gpurun_out/ncu_$tag.log 2>&1
  ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > /tmp/ncu/$tag.csv 2>/dev/null
  python scripts/ncu_compact.py /tmp/ncu/$tag.csv > gpurun_out/ncu_$tag.csv 2>> gpurun_out/ncu_$tag.log
done
ls -la /tmp/ncu gpurun_out | head -60
du -sh gpurun_out
