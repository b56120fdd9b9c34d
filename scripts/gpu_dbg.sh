#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; : > $O/dbg.txt
HOP_VAR=1 MGCR_HOPPING_TMA_ROWS=0 timeout 60 python scripts/hop_bench.py 2 64x64x128 40x33x130 2>&1 | tail -3 >> $O/dbg.txt
cat $O/dbg.txt
if grep -q "exact=True" $O/dbg.txt; then bash scripts/gpu_call10.sh; fi
