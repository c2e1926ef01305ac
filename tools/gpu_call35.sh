#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_elementwise_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/c35_per_launch.txt 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['by_kind_ms'])"
) > gpurun_out/c35.log 2>&1
cat gpurun_out/c35.log
