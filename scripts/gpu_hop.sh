#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; : > $O/hop4.txt
timeout 600 python -m pytest tests/test_gpu_krylov.py tests/test_gpu_mg.py -m gpu -x -q 2>&1 | tail -3
for w in gcr3d_256 gcr3d_512 gcr2d_4096; do for tx in 64 256; do echo "HL_TX=$tx $w" >> $O/hop4.txt; MGCR_HL_TX=$tx timeout 200 python bench.py --workload $w --operator stencil --steps 1 --warmup 1 --no-cpu-baseline --max-iter 40 2>&1 | tail -1 >> $O/hop4.txt; done; done
python - <<'PY'
import json
cur=None
for ln in open('gpurun_out/hop4.txt'):
    ln=ln.strip()
    if not ln.startswith('{'): cur=ln; continue
    try:
        j=json.loads(ln); k=j['kernels']['hopping_dirac']; print(cur,'hopping %.1f us %.0f GB/s'%(k['ms_per_launch']*1e3,k['GBps']))
    except Exception as e: print(cur,'ERR',ln[:300])
PY
