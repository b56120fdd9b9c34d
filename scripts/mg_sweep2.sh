#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/mg_sweep2.txt; : > $O
run() { echo "CFG $1 $2 restart=$3" >> $O; timeout 300 python bench.py --workload $1 --operator stencil --steps 1 --warmup 1 --no-cpu-baseline --mg "$2" --restart $3 2>&1 | tail -1 | python -c "
import sys,json
try:
    j=json.loads(sys.stdin.read()); print('   value %.4f iters %d setup %.2f res %.2e launches %d alloc_ms %.0f'%(j['value'],j['iterations'],j['mg_setup_seconds'],j['final_true_rel_residual'],j['gpu_launches'],j['host_side'].get('host_alloc',{}).get('ms',0)))
except Exception as e: print('   FAILED',e)" >> $O; }
for r in 3 5 8; do for cm in 2 3 4; do for sm in 1 2 3; do run mg3d_256 "{\"coarse\": [0,10,$cm,0.01], \"smooth\": [0,4,$sm,1e-8]}" $r; done; done; done
run mg3d_256 '{"coarse": [0,10,4,0.1], "smooth": [0,4,2,1e-8]}' 5
run mg3d_256 '{"coarse": [0,4,4,0.01], "smooth": [0,2,2,1e-8]}' 5
run mg3d_512 '{}' 10
run mg3d_512 '{"coarse": [0,10,4,0.01], "smooth": [0,4,2,1e-8]}' 5
run mg3d_512 '{"coarse": [0,10,2,0.01], "smooth": [0,4,2,1e-8]}' 5
run mg3d_512 '{"coarse": [0,10,3,0.01], "smooth": [0,4,2,1e-8]}' 5
run mg3d_512 '{"coarse": [0,10,4,0.01], "smooth": [0,4,3,1e-8]}' 5
cat $O
