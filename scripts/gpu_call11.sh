#!/bin/bash
# sliced block-CSR: parity, then the three MG workloads
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
summ() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); print('$1', j['config']['workload'],'value',j['value'],'iters',j['iterations'],'res',j['final_true_rel_residual'],'setup',j['mg_setup_seconds'],'e2e',j['e2e']['value'],'roofline',j['roofline']['kernel'],round(j['roofline']['frac'],3),'spmv',j['spmv']['kernel'],round(j['spmv']['frac'],3))
for k,v in sorted(j['kernels'].items(), key=lambda kv:-kv[1]['share'])[:9]: print('   %-20s share %.3f  %8.1f us  %6.0f GB/s  x%d'%(k,v['share'],v['ms_per_launch']*1e3,v['GBps'] or 0,v['launches']))
" 2>&1 | tail -12; }
timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/b512.json 2>$O/b512.err; summ $O/b512.json

timeout 300 python bench.py --workload mg3d_aniso_512 --steps 2 --warmup 1 --no-cpu-baseline > $O/baniso512.json 2>$O/baniso512.err; summ $O/baniso512.json; tail -3 $O/baniso512.err
