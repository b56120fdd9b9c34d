#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
: > $O/mg_repeat2.txt
for op in stencil csr; do timeout 300 python bench.py --workload mg3d_256 --operator $op --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 >> $O/mg_repeat2.txt; done
MGCR_SMALL_GCR_N=0 timeout 300 python bench.py --workload mg3d_256 --operator stencil --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 >> $O/mg_repeat2.txt
python - <<'PY'
import json
for ln in open('gpurun_out/mg_repeat2.txt'):
    try:
        j=json.loads(ln); ks=j['kernels']; tot=sum(v['ms_per_launch']*v['launches'] for v in ks.values())
        print(j['config']['operator'],'value %.4f iters %d launches %d kernel-sum %.1f ms setup %.2f'%(j['value'],j['iterations'],j['gpu_launches'],tot,j['mg_setup_seconds']), j['host_side'])
    except Exception as e: print('ERR',ln[:300])
PY
