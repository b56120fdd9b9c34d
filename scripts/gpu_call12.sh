#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; : > $O/bcsr.txt
for st in 7 13; do MGCR_BLOCKCSR_STAGES=$st timeout 120 python scripts/bcsr_bench.py 4 96 >> $O/bcsr.txt 2>&1; done
for st in 7 14; do MGCR_BLOCKCSR_STAGES=$st timeout 120 python scripts/bcsr_bench.py 2 128 >> $O/bcsr.txt 2>&1; done
timeout 120 python scripts/bcsr_bench.py 8 64 >> $O/bcsr.txt 2>&1
cat $O/bcsr.txt
