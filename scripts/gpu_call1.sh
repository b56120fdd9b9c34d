#!/bin/bash
# one gpurun call: parity tests, benches, knob experiments.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 > $O/bench_gcr2d.json 2> $O/bench_gcr2d.err; tail -c 600 $O/bench_gcr2d.err
timeout 600 python bench.py --workload mg3d_256 --steps 1 --warmup 1 --no-cpu-baseline > $O/bench_mg256_small.json 2> $O/bench_mg256_small.err; tail -c 600 $O/bench_mg256_small.err
MGCR_SMALL_GCR_N=0 timeout 600 python bench.py --workload mg3d_256 --steps 1 --warmup 1 --no-cpu-baseline > $O/bench_mg256_nosmall.json 2> $O/bench_mg256_nosmall.err
timeout 300 python bench.py --workload gcr3d_256 --operator stencil --steps 1 --warmup 1 --no-cpu-baseline > $O/bench_gcr3d256.json 2> $O/bench_gcr3d256.err
# knob experiments (stencil operator: no CSR build), 200 iterations each
X="python bench.py --workload gcr2d_4096 --operator stencil --steps 1 --warmup 1 --no-cpu-baseline --max-iter 200"
: > $O/knobs.txt
for pad in 0 64 1040 4112; do echo "RING_PAD=$pad" >> $O/knobs.txt; MGCR_RING_PAD=$pad timeout 200 $X >> $O/knobs.txt 2>&1; done
for g in 2 8; do echo "GRID_PER_SM=$g" >> $O/knobs.txt; MGCR_GRID_PER_SM=$g timeout 200 $X >> $O/knobs.txt 2>&1; done
for u in 1 2 4; do echo "DOT_U=$u" >> $O/knobs.txt; MGCR_DOT_U=$u timeout 200 $X >> $O/knobs.txt 2>&1; done
for m in 2 3; do echo "UPD_MINB=$m" >> $O/knobs.txt; MGCR_UPD_MINB=$m timeout 200 $X >> $O/knobs.txt 2>&1; done
echo done
