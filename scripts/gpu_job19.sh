#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python scripts/kernel_zoo.py --only "adamw" 2>&1 | tail -6
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n "$@"; }
run 2 --workload pretrain --steps 3 --warmup 3 --no-cpu-baseline --zero1 2>gpurun_out/z1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('zero1', round(d['value']), d['final_loss'], d['config'])" || tail -20 gpurun_out/z1.err
run 2 --workload pretrain --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('allreduce', round(d['value']), d['final_loss'])"
