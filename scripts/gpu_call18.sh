#!/bin/bash
# ncu --set full of the block-CSR ring kernel inside the default workload, then the full GPU suite on the final tree
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --max-iter 2"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_blockcsr_ring' --launch-skip 30 -c 3 -f -o $O/r01b_mg512_blockcsr $CMD > $O/ncu_c.log 2>&1; tail -1 $O/ncu_c.log
timeout 300 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
