#!/bin/bash
# experiment: ring-buffer stride padding and grid size vs kernel bandwidth (gcr2d_4096, 45 iterations)
for pad in 0 264 4104 65544; do for g in 4 8; do
echo "pad=$pad grid_per_sm=$g"
MGCR_RING_PAD=$pad MGCR_GRID_PER_SM=$g python bench.py --steps 1 --warmup 1 --max-iter 45 --no-cpu-baseline | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  ms/step %.2f'%j['ms_per_step'], {k:(round(v['GBps']),round(v['ms_per_launch'],3)) for k,v in j['kernels'].items()})"
done; done
