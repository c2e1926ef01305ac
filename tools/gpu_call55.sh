#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_elementwise_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(d['ms_per_step'], d['e2e']['ms_per_step'], r['frac'], r['by_kind_ms'])"
) > gpurun_out/c55.log 2>&1
cat gpurun_out/c55.log
