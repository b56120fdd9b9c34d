#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
: > $O/hop.txt
for hk in 0 1; do for w in gcr3d_256 gcr2d_4096; do echo "HOPPING_KERNEL=$hk $w" >> $O/hop.txt; MGCR_HOPPING_KERNEL=$hk timeout 200 python bench.py --workload $w --operator stencil --steps 1 --warmup 1 --no-cpu-baseline --max-iter 100 2>&1 | tail -1 >> $O/hop.txt; done; done
python - <<'PY'
import json
cur=None
for ln in open('gpurun_out/hop.txt'):
    ln=ln.strip()
    if not ln.startswith('{'): cur=ln; continue
    try:
        j=json.loads(ln); k=j['kernels']['hopping_dirac']; print(cur,'value %.4f hopping %.1f us %.0f GB/s'%(j['value'],k['ms_per_launch']*1e3,k['GBps']))
    except Exception as e: print(cur,'ERR',ln[:300])
PY
O2=$O/mg_sweep3.txt; : > $O2
run() { echo "CFG $1 $2 restart=$3" >> $O2; timeout 300 python bench.py --workload $1 --operator stencil --steps 1 --warmup 1 --no-cpu-baseline --mg "$2" --restart $3 2>&1 | tail -1 | python -c "
import sys,json
try:
    j=json.loads(sys.stdin.read()); ks=j['kernels']
    print('   value %.4f iters %d setup %.2f res %.2e launches %d'%(j['value'],j['iterations'],j['mg_setup_seconds'],j['final_true_rel_residual'],j['gpu_launches']), ' '.join('%s:%.2f/%.0f'%(k,v['share'],v['GBps'] or 0) for k,v in sorted(ks.items(), key=lambda kv:-kv[1]['share'])[:6]))
except Exception as e: print('   FAILED',e)" >> $O2; }
run mg3d_512 '{"coarse": [0,10,2,0.01], "smooth": [0,4,2,1e-8]}' 3
run mg3d_512 '{"coarse": [0,10,3,0.01], "smooth": [0,4,2,1e-8]}' 3
run mg3d_512 '{"coarse": [0,10,2,0.01], "smooth": [0,4,3,1e-8]}' 3
run mg3d_256 '{"coarse": [0,10,3,0.01], "smooth": [0,4,2,1e-8]}' 3
run mg3d_256 '{"coarse": [0,10,3,0.01], "smooth": [0,4,2,1e-8]}' 2
run mg3d_256 '{"coarse": [0,10,3,0.01], "smooth": [0,4,2,1e-8]}' 4
cat $O2
