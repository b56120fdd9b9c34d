#!/bin/bash
# last check of the round on 2 GPUs: peer-memory exchanges with spin guards, residual-form prefetch in the TMA stencil
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k "2-default-var" > $O/pytest_dist2c.log 2>&1; echo "pytest rc=$?" >> $O/pytest_dist2c.log
tail -4 $O/pytest_dist2c.log
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > $O/bench_p2p_n2b.json 2>$O/bench_p2p_n2b.err
python -c "
import json
j=json.loads(open('$O/bench_p2p_n2b.json').read().strip().splitlines()[-1]); print('N=2 value',j['value'],'iters',j['iterations'],'res',j['final_true_rel_residual'])
for k,v in sorted(j['kernels'].items(), key=lambda kv:-kv[1]['share'])[:6]: print('   %-20s share %.3f  %8.1f us x%d %s'%(k,v['share'],v['ms_per_launch']*1e3,v['launches'],v['GBps']))
" 2>&1 | tail -8
