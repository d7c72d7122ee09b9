#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
for sel in adamw layernorm_bwd colsum; do
  tag=$(echo "$sel" | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:vlk:: -c 40 \
     -o /tmp/ncu/$tag -f python scripts/kernel_zoo.py --once --only "$sel" > gpurun_out/ncu_$tag.log 2>&1
  ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > /tmp/ncu/$tag.csv 2>/dev/null
  python scripts/ncu_compact.py /tmp/ncu/$tag.csv > gpurun_out/ncu_$tag.csv 2>> gpurun_out/ncu_$tag.log
done
ls -la gpurun_out/ncu_*.csv
