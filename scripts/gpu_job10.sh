#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
for k in flash_fwd_kernel flash_bwd_dkv_kernel flash_bwd_dq_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -c 1 \
     -o /tmp/ncu/$k -f python scripts/kernel_zoo.py --once --only "pretrain B=16 H=12 T=1024" > gpurun_out/ncu_$k.log 2>&1
  ncu -i /tmp/ncu/$k.ncu-rep --page source --csv > /tmp/ncu/$k.csv 2>/dev/null
  echo "=== $k"; python scripts/ncu_source_summary.py /tmp/ncu/$k.csv 28 | tee gpurun_out/src_$k.txt
  ncu -i /tmp/ncu/$k.ncu-rep --page raw --csv > /tmp/ncu/$k.raw.csv 2>/dev/null; python scripts/ncu_compact.py /tmp/ncu/$k.raw.csv > gpurun_out/ncu_$k.csv
done
