#!/bin/bash
# round-1 record: full GPU suite, the bench lines of every workload, ncu launch list + full captures of the default workload
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
summ() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); print('$1', j['config']['workload'],'value',j['value'],'iters',j['iterations'],'res',j['final_true_rel_residual'],'setup',j['mg_setup_seconds'],'e2e',j['e2e']['value'],'roofline',j['roofline']['kernel'],round(j['roofline']['frac'],3),'spmv',j['spmv']['kernel'],round(j['spmv']['frac'],3), 'cpu', j.get('cpu_baseline',{}).get('value'))
for k,v in sorted(j['kernels'].items(), key=lambda kv:-kv[1]['share'])[:9]: print('   %-20s share %.3f  %8.1f us  %6.0f GB/s  x%d'%(k,v['share'],v['ms_per_launch']*1e3,v['GBps'] or 0,v['launches']))
" 2>&1 | tail -12; }
timeout 600 python bench.py > $O/bench_default_n1.json 2>$O/bench_default_n1.err; summ $O/bench_default_n1.json
timeout 300 python bench.py --workload mg3d_aniso --steps 2 --warmup 1 > $O/bench_aniso_n1.json 2>$O/bench_aniso_n1.err; summ $O/bench_aniso_n1.json
timeout 300 python bench.py --workload mg3d_256 --steps 3 --warmup 3 --operator stencil --no-cpu-baseline > $O/bench_mg256_stencil.json 2>$O/bench_mg256.err; summ $O/bench_mg256_stencil.json
timeout 300 python bench.py --workload gcr2d_4096 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_gcr2d.json 2>$O/bench_gcr2d.err; summ $O/bench_gcr2d.json
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --max-iter 2"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $O/r01b_launches_mg3d_512.csv $CMD > $O/ncu_a.log 2>&1; tail -1 $O/ncu_a.log
ncu --set full --clock-control none --import-source on -k regex:'k_hopping_tma|k_blockcsr_ring' --launch-skip 20 -c 6 -f -o $O/r01b_mg512_ops $CMD > $O/ncu_b.log 2>&1; tail -1 $O/ncu_b.log
ls -la $O/r01b*
