#!/bin/bash
# 2 GPUs: peer-memory halo / all-reduce against the single-GPU results, then the default bench at N=2
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
export MGCR_VERBOSE=1
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k "2-" > $O/pytest_dist2.log 2>&1; echo "pytest rc=$?" >> $O/pytest_dist2.log
tail -15 $O/pytest_dist2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_p2p_n2.json 2>$O/bench_p2p_n2.err
python -c "
import json
j=json.loads(open('$O/bench_p2p_n2.json').read().strip().splitlines()[-1]); print('N=2 value',j['value'],'iters',j['iterations'],'res',j['final_true_rel_residual'])
for k,v in sorted(j['kernels'].items(), key=lambda kv:-kv[1]['share']): print('   %-20s share %.3f  %8.1f us x%d'%(k,v['share'],v['ms_per_launch']*1e3,v['launches']))
" 2>&1 | tail -20; tail -3 $O/bench_p2p_n2.err
