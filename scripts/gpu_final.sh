#!/bin/bash
# end-of-round check: GPU tests, smoke(), the default bench line and the other workloads
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; wc -l < gpurun_out/bench_default.json; cut -c1-260 gpurun_out/bench_default.json
for w in qformer xattn; do python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2>/dev/null; cut -c1-150 gpurun_out/bench_$w.json; done
python bench.py --workload pretrain --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pretrain.json 2>/dev/null; cut -c1-150 gpurun_out/bench_pretrain.json
