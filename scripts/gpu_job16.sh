#!/bin/bash
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_linear.json 2> gpurun_out/bench_linear.err || tail -5 gpurun_out/bench_linear.err
python bench.py --workload pretrain --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pretrain.json 2> gpurun_out/bench_pretrain.err || tail -5 gpurun_out/bench_pretrain.err
python - <<'PY'
import json
for w in ("linear","pretrain"):
    d=json.loads(open(f"gpurun_out/bench_{w}.json").read())
    print(w, round(d["value"],1), json.dumps(d["roofline"]))
PY
python -m pytest tests -m gpu -x -q -W default 2>&1 | grep -i -A3 "warn" | head -30
