#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/b512.json 2>$O/b512.err; python -c "
import json
j=json.loads(open('$O/b512.json').read().strip().splitlines()[-1]); print('mg3d_512 value',j['value'],'iters',j['iterations'],'setup',j['mg_setup_seconds'],'e2e',j['e2e']['value'],'roofline',j['roofline']['kernel'],j['roofline']['frac'],j['roofline']['traffic'])"
