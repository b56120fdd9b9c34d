#!/bin/bash
# new variable-coefficient tests + TMA stencil sweep
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_var.py tests/test_gpu_dropin.py tests/test_gpu_krylov.py -m gpu -q > $O/pytest_var.log 2>&1; echo "pytest rc=$?" >> $O/pytest_var.log
tail -5 $O/pytest_var.log
: > $O/hop5.txt
L="40x33x130 256x256x256 512x512x512 64x512x512"
timeout 120 python scripts/hop_bench.py 1 $L >> $O/hop5.txt 2>&1
for st in 4 6 8 10; do MGCR_HOP_STAGES=$st timeout 120 python scripts/hop_bench.py 2 $L >> $O/hop5.txt 2>&1; done
for zc in 16 32 128; do MGCR_HOP_ZC=$zc timeout 120 python scripts/hop_bench.py 2 $L >> $O/hop5.txt 2>&1; done
MGCR_HOP_TILE=1 timeout 120 python scripts/hop_bench.py 2 $L >> $O/hop5.txt 2>&1
MGCR_HOP_TILE=1 MGCR_HOP_STAGES=10 timeout 120 python scripts/hop_bench.py 2 $L >> $O/hop5.txt 2>&1
cat $O/hop5.txt
