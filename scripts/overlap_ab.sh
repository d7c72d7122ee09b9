#!/bin/bash
# A/B of the split-backward gradient-exchange overlap on N GPUs of one box (default 4): cross-attention captioner and
# GPT-2 pretraining, overlap on (default) vs VLK_NO_OVERLAP=1.  Usage: gpurun --gpus 4 -- bash scripts/overlap_ab.sh 4
N=${1:-4}
mkdir -p gpurun_out/overlap_ab
for w in xattn pretrain; do
  for mode in ov noov; do
    if [ $mode = noov ]; then export VLK_NO_OVERLAP=1; else unset VLK_NO_OVERLAP; fi
    steps=20; [ $w = pretrain ] && steps=4
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $N --workload $w --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/overlap_ab/${w}_${mode}_n$N.json 2> gpurun_out/overlap_ab/${w}_${mode}_n$N.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/overlap_ab/${w}_${mode}_n$N.json"))
    print("$w", "$mode", "n=$N", round(d["value"], 1), d["unit"], round(d["ms_per_step"], 3), "ms", "overlap_comm", d["config"]["overlap_comm"], "dp_check", (d.get("dp_check") or {}).get("ok_all_ranks"))
except Exception as e:
    print("$w $mode failed", e)
PY
  done
done
