#!/bin/bash
# One `ncu --set full` capture of every attention kernel at its bench shape (after the same command ran clean without ncu):
# the compact metric table and the source-page stall summaries of the three streaming kernels, processed on the GPU box.
mkdir -p gpurun_out/r02 /tmp/ncu
O=gpurun_out/r02
timeout 120 python scripts/kernel_zoo.py --once --only "attention" > $O/attn_plain.log 2>&1 && \
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:vlk:: -c 60 \
   -o /tmp/ncu/attention -f python scripts/kernel_zoo.py --once --only "attention" > $O/ncu_attention.log 2>&1
ncu -i /tmp/ncu/attention.ncu-rep --page raw --csv > /tmp/ncu/attention.csv 2>/dev/null
python scripts/ncu_compact.py /tmp/ncu/attention.csv > $O/ncu_attention.csv 2>> $O/ncu_attention.log
for k in flash_bwd_dkv flash_bwd_dq flash_fwd; do
  ncu -i /tmp/ncu/attention.ncu-rep --page source --csv --kernel-name-base demangled -k regex:$k -c 1 > /tmp/ncu/src_$k.csv 2>/dev/null
  python scripts/ncu_source_top.py /tmp/ncu/src_$k.csv 25 > $O/ncu_src_$k.txt 2>&1
done
ls -la $O /tmp/ncu | head -30
