#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
run 2 --workload xattn --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ov_xattn_n2.json 2> gpurun_out/ov_xattn_n2.err
VLK_NO_OVERLAP=1 run 2 --workload xattn --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/noov_xattn_n2.json 2> gpurun_out/noov_xattn_n2.err
run 2 --workload pretrain --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ov_pretrain_n2.json 2> gpurun_out/ov_pretrain_n2.err
VLK_NO_OVERLAP=1 run 2 --workload pretrain --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/noov_pretrain_n2.json 2> gpurun_out/noov_pretrain_n2.err
run 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ov_linear_n2.json 2> gpurun_out/ov_linear_n2.err
for f in gpurun_out/*ov_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().split('\n')[-1]); print(round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'loss', d.get('final_loss'), d['config'].get('parallelism'), d['config'].get('overlap_comm'))"; done
tail -n 5 gpurun_out/ov_xattn_n2.err gpurun_out/ov_pretrain_n2.err
