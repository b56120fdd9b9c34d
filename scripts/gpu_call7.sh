#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
show() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); print('$1 value',j['value'],'iters',j['iterations'],'res',j['final_true_rel_residual'],'e2e',j['e2e']['value'])
for k,v in sorted(j['kernels'].items(), key=lambda kv:-kv[1]['share']): print('   %-16s %8.1f us x %6d share %.3f %s'%(k,v['ms_per_launch']*1e3,v['launches'],v['share'],v['GBps']))
"; }
timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/b512.json 2>$O/b512.err; show $O/b512.json
timeout 300 python bench.py --workload mg3d_256 --operator csr --steps 2 --warmup 2 > $O/b256csr.json 2>$O/b256csr.err; show $O/b256csr.json
timeout 300 python bench.py --workload gcr2d_4096 --steps 2 --warmup 3 > $O/b2d.json 2>$O/b2d.err; show $O/b2d.json
