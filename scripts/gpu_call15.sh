#!/bin/bash
# 8 GPUs: default workload (configs[3]) and the anisotropic 1024x512x512 workload (configs[4])
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { # name, extra args
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline $2 > $O/$1.json 2>$O/$1.err
  python -c "
import json
j=json.loads(open('$O/$1.json').read().strip().splitlines()[-1]); print('$1 value',j['value'],'iters',j['iterations'],'res',j['final_true_rel_residual'],'setup',j['mg_setup_seconds'],'e2e',j['e2e']['value'])
for k,v in sorted(j['kernels'].items(), key=lambda kv:-kv[1]['share']): print('   %-20s share %.3f  %8.1f us x%d %s'%(k,v['share'],v['ms_per_launch']*1e3,v['launches'], '%.0f GB/s'%v['GBps'] if v['GBps'] else ''))
" 2>&1 | tail -18; tail -2 $O/$1.err; }
run bench_default_n8 "" 29621
run bench_aniso_n8 "--workload mg3d_aniso" 29622
