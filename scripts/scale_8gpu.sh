#!/bin/bash
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n "$@"; }
run 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_linear_n8.json 2> gpurun_out/scale_linear_n8.err
run 8 --workload pretrain --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/scale_pretrain_n8.json 2> gpurun_out/scale_pretrain_n8.err
run 8 --workload xattn --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_xattn_n8.json 2> gpurun_out/scale_xattn_n8.err
for f in gpurun_out/scale_*_n8.json; do echo $f; wc -l < $f; cut -c1-170 $f; done
tail -n 3 gpurun_out/scale_linear_n8.err
