#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
nproc > $O/host.txt; cat /proc/loadavg >> $O/host.txt
: > $O/mg_repeat.txt
for i in 1 2 3; do timeout 200 python bench.py --workload mg3d_256 --operator stencil --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 >> $O/mg_repeat.txt; done
for i in 1 2; do timeout 200 python bench.py --workload mg3d_256 --operator stencil --steps 2 --warmup 1 --no-cpu-baseline --mg '{"coarse": [0,10,1,0.01]}' 2>&1 | tail -1 >> $O/mg_repeat.txt; done

cat /proc/loadavg >> $O/host.txt
python - <<'PY'
import json
for ln in open('gpurun_out/mg_repeat.txt'):
    try:
        j=json.loads(ln); ks=j['kernels']; tot=sum(v['ms_per_launch']*v['launches'] for v in ks.values())
        print('value %.4f iters %d launches %d kernel-sum %.1f ms setup %.2f'%(j['value'],j['iterations'],j['gpu_launches'],tot,j['mg_setup_seconds']))
    except Exception as e: print('ERR',ln[:200])
PY
cat $O/host.txt
