#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
# source-level profile of the CLIP attention kernel and of the residual GEMM (reports stay on the box; CSV pages travel)
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:attn_fwd_tcgen05_v4 -c 1 \
   -o /tmp/ncu/attn_v4 -f python scripts/kernel_zoo.py --once --only "attention fwd CLIP" > gpurun_out/ncu_attn_v4.log 2>&1
ncu -i /tmp/ncu/attn_v4.ncu-rep --page source --csv > gpurun_out/attn_v4_source.csv 2>/dev/null
ncu -i /tmp/ncu/attn_v4.ncu-rep --page details --csv > gpurun_out/attn_v4_details.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gemm_bf16_2cta -c 1 \
   -o /tmp/ncu/gemm_res -f python scripts/kernel_zoo.py --once --only "gemm CLIP out_proj" > gpurun_out/ncu_gemm_res.log 2>&1
ncu -i /tmp/ncu/gemm_res.ncu-rep --page source --csv > gpurun_out/gemm_res_source.csv 2>/dev/null
ncu -i /tmp/ncu/gemm_res.ncu-rep --page details --csv > gpurun_out/gemm_res_details.csv 2>/dev/null
ls -la gpurun_out/*.csv; du -sh gpurun_out
