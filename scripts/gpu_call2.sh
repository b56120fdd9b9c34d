#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
X="python bench.py --workload gcr2d_4096 --operator stencil --steps 1 --warmup 1 --no-cpu-baseline --max-iter 200"
: > $O/knobs2.txt
for t in 1 0; do echo "DOT_TMA=$t" >> $O/knobs2.txt; MGCR_DOT_TMA=$t timeout 200 $X >> $O/knobs2.txt 2>&1; done
timeout 600 python bench.py --steps 2 --warmup 3 > $O/bench_gcr2d.json 2> $O/bench_gcr2d.err; tail -c 300 $O/bench_gcr2d.err
echo done
