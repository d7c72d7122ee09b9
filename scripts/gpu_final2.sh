#!/bin/bash
# end-of-round evidence: GPU tests, smoke(), graph-timed kernel zoo, the default bench line (all workloads)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python scripts/kernel_zoo.py > $O/kernel_zoo_graph_timed.log 2>&1; grep -c "us" $O/kernel_zoo_graph_timed.log
timeout 600 python bench.py > $O/bench_n1_default.json 2> $O/bench_n1_default.err; wc -l < $O/bench_n1_default.json; cut -c1-300 $O/bench_n1_default.json
