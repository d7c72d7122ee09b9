#!/bin/bash
# multi-GPU scaling sanity on one box: weak scaling of the caption steps, strong scaling of the pretraining step
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
run 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_linear_n4.json 2> gpurun_out/scale_linear_n4.err
run 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_linear_n2.json 2> gpurun_out/scale_linear_n2.err
run 4 --workload xattn --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_xattn_n4.json 2> gpurun_out/scale_xattn_n4.err
run 4 --workload pretrain --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_pretrain_n4.json 2> gpurun_out/scale_pretrain_n4.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_linear_n1.json 2> gpurun_out/scale_linear_n1.err
for f in gpurun_out/scale_*.json; do echo $f; cut -c1-160 $f; done
tail -3 gpurun_out/scale_*.err | tail -30
