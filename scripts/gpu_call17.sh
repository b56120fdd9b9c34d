#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k "2-default-unit" > $O/pytest_dist2b.log 2>&1; echo "pytest rc=$?" >> $O/pytest_dist2b.log
tail -25 $O/pytest_dist2b.log
