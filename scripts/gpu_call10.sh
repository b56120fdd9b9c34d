#!/bin/bash
# full GPU suite with the TMA stencil as default, variable-coefficient stencil rates, default bench, anisotropic workloads on one GPU
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
: > $O/hop6.txt
L="256x256x256 512x256x512"
HOP_VAR=1 timeout 200 python scripts/hop_bench.py 1 $L >> $O/hop6.txt 2>&1
for st in 3 4; do HOP_VAR=1 MGCR_HOP_STAGES=$st timeout 200 python scripts/hop_bench.py 2 $L >> $O/hop6.txt 2>&1; done
cat $O/hop6.txt
summ() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); print('$1', j['config']['workload'],'value',j['value'],'iters',j['iterations'],'res',j['final_true_rel_residual'],'setup',j['mg_setup_seconds'],'e2e',j['e2e']['value'],'roofline',j['roofline']['kernel'],round(j['roofline']['frac'],3),'spmv',j['spmv']['kernel'],round(j['spmv']['frac'],3))
for k,v in sorted(j['kernels'].items(), key=lambda kv:-kv[1]['share'])[:9]: print('   %-20s share %.3f  %8.1f us  %6.0f GB/s  x%d'%(k,v['share'],v['ms_per_launch']*1e3,v['GBps'] or 0,v['launches']))
" 2>&1 | tail -12; }
timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/b512.json 2>$O/b512.err; summ $O/b512.json
timeout 300 python bench.py --workload mg3d_aniso_512 --steps 2 --warmup 1 --no-cpu-baseline > $O/baniso512.json 2>$O/baniso512.err; summ $O/baniso512.json; tail -3 $O/baniso512.err
timeout 400 python bench.py --workload mg3d_aniso --steps 1 --warmup 1 --no-cpu-baseline > $O/baniso.json 2>$O/baniso.err; summ $O/baniso.json; tail -3 $O/baniso.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
