#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --max-iter 2"
$CMD > $O/ncu_plain.log 2>&1 || { echo plain failed; tail -5 $O/ncu_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $O/r01_launches_mg3d_512.csv $CMD > $O/ncu_a.log 2>&1; tail -1 $O/ncu_a.log
ncu --set full --clock-control none --import-source on -k regex:'k_gcr_update_xr|k_gcr_update_p|k_hopping_l1|k_gcr_init|k_gcr_dot_hist' --launch-skip 40 -c 12 -f -o $O/r01_mg512_krylov $CMD > $O/ncu_b.log 2>&1; tail -1 $O/ncu_b.log
ncu --set full --clock-control none --import-source on -k regex:'k_restrict_warp|k_prolong' -c 4 -f -o $O/r01_mg512_transfer $CMD > $O/ncu_c.log 2>&1; tail -1 $O/ncu_c.log
ncu --set full --clock-control none --import-source on -k regex:'k_blockcsr_apply_ne' --launch-skip 30 -c 3 -f -o $O/r01_mg512_blockcsr $CMD > $O/ncu_d.log 2>&1; tail -1 $O/ncu_d.log
ls -la $O/*.ncu-rep $O/r01_launches_mg3d_512.csv
