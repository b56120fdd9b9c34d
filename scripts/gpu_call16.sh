#!/bin/bash
# final single-GPU check of the tree: full suite, smoke, anisotropic workload after the memory fixes, default bench
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
summ() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); print('$1', j['config']['workload'],'value',j['value'],'iters',j['iterations'],'res',j['final_true_rel_residual'],'setup',j['mg_setup_seconds'],'e2e',j['e2e']['value'],'roofline',j['roofline']['kernel'],round(j['roofline']['frac'],3),'host',j['host_side'])
for k,v in sorted(j['kernels'].items(), key=lambda kv:-kv[1]['share'])[:9]: print('   %-20s share %.3f  %8.1f us  %6.0f GB/s  x%d'%(k,v['share'],v['ms_per_launch']*1e3,v['GBps'] or 0,v['launches']))
" 2>&1 | tail -12; }
timeout 300 python bench.py --workload mg3d_aniso --steps 1 --warmup 1 --no-cpu-baseline > $O/bench_aniso_n1b.json 2>$O/bench_aniso_n1b.err; summ $O/bench_aniso_n1b.json; tail -2 $O/bench_aniso_n1b.err
timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/bench_default_n1b.json 2>$O/bench_default_n1b.err; summ $O/bench_default_n1b.json
