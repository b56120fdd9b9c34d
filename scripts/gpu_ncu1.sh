#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
CMD="python bench.py --workload mg3d_256 --operator stencil --steps 1 --warmup 1 --no-cpu-baseline --max-iter 2"
$CMD > $O/ncu_plain.log 2>&1 || { echo plain failed; tail -5 $O/ncu_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k_hopping_l1|k_blockcsr_apply_ne|k_gcr_dot_hist_tma|k_restrict|k_prolong|k_gcr_update_p|k_gcr_update_xr' --launch-skip 400 --launch-count 24 -f -o $O/r01_mg256 $CMD > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log
ls -la $O/r01_mg256.ncu-rep
