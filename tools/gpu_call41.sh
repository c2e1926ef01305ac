#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( for m in 0 1; do RD_WG_SW128=$m timeout 300 python tools/check_wgrad_sw128.py s2 s3 ds3 up1 up2 up3 l2 2>&1 | tail -12; done
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 200 python tools/bench_fprop.py l1 bn 2>&1 | head -8
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/c41_per_launch.txt 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['by_kind_ms'])"
timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
) > gpurun_out/c41.log 2>&1
cat gpurun_out/c41.log
