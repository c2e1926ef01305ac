#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( for m in 0 1; do RD_WG_SW128=$m timeout 300 python tools/check_wgrad_sw128.py l1 l2 l3 2>&1 | tail -12; done
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['by_kind_ms'])"
) > gpurun_out/c34.log 2>&1
cat gpurun_out/c34.log
