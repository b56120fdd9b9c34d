#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
: > $O/small.txt
for ps in 1 2; do for rows in 256 512 1024 2048; do echo "PER_SM=$ps ROWS=$rows" >> $O/small.txt; MGCR_SMALL_GRID_PER_SM=$ps MGCR_SMALL_ROWS_PER_CTA=$rows timeout 200 python bench.py --workload mg3d_256 --operator stencil --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 >> $O/small.txt; done; done
python - <<'PY'
import json
cur=None
for ln in open('gpurun_out/small.txt'):
    ln=ln.strip()
    if not ln.startswith('{'): cur=ln; continue
    try:
        j=json.loads(ln); k=j['kernels']['gcr_small']; print(cur,'value %.4f iters %d gcr_small %.1f us x %d'%(j['value'],j['iterations'],k['ms_per_launch']*1e3,k['launches']))
    except Exception as e: print(cur,'ERR',ln[:300])
PY
timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/bench_n1_fused.json 2>$O/bench_n1_fused.err; python -c "
import json
j=json.loads(open('gpurun_out/bench_n1_fused.json').read().strip().splitlines()[-1]); print('mg3d_512 N=1 value',j['value'],'iters',j['iterations'],'res',j['final_true_rel_residual'])
for k,v in sorted(j['kernels'].items(), key=lambda kv:-kv[1]['share']): print('   %-16s %8.1f us x %6d share %.3f %s'%(k,v['ms_per_launch']*1e3,v['launches'],v['share'],v['GBps']))
"
