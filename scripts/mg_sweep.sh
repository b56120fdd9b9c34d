#!/bin/bash
# parameter sweep of the 256^3 3-level MG-GCR (matrix-free operator): time-to-solution vs cycle parameters
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/mg_sweep.txt; : > $O
run() { echo "CFG $1 restart=$2" >> $O; timeout 120 python bench.py --workload mg3d_256 --operator stencil --steps 1 --warmup 1 --no-cpu-baseline --mg "$1" --restart $2 2>&1 | tail -1 | python -c "
import sys,json
try:
    j=json.loads(sys.stdin.read()); print('   value %.4f iters %d setup %.2f res %.2e launches %d'%(j['value'],j['iterations'],j['mg_setup_seconds'],j['final_true_rel_residual'],j['gpu_launches']))
except Exception as e: print('   FAILED',e)" >> $O; }
run '{}' 10
for cm in 1 2 4 8; do run "{\"coarse\": [0,10,$cm,0.01]}" 10; done
for sm in 0 1 2; do run "{\"smooth\": [0,4,$sm,1e-8]}" 10; done
for sm in 1 2; do for cm in 2 4; do run "{\"smooth\": [0,4,$sm,1e-8], \"coarse\": [0,10,$cm,0.01]}" 10; done; done
run '{"n_eigen": [8,8]}' 10
run '{"n_eigen": [8,8], "coarse": [0,10,4,0.01], "smooth": [0,4,2,1e-8]}' 10
run '{"n_eigen": [8,4], "coarse": [0,10,4,0.01], "smooth": [0,4,2,1e-8]}' 10
run '{"n_eigen": [2,2]}' 10
run '{"n_eigen": [2,2], "coarse": [0,10,4,0.01], "smooth": [0,4,2,1e-8]}' 10
run '{"coarse": [0,10,4,0.01], "smooth": [0,4,2,1e-8]}' 5
run '{"coarse": [0,10,4,0.01], "smooth": [0,4,2,1e-8]}' 20
run '{"subs": [8,4], "n_eigen": [8,4]}' 10
run '{"subs": [8,4], "n_eigen": [8,4], "coarse": [0,10,4,0.01], "smooth": [0,4,2,1e-8]}' 10
cat $O
