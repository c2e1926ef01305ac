#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_tuned_tiles_gpu.py tests/test_model_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 200 python tools/bench_fprop.py stem 2>&1 | head -3
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(d['ms_per_step'], d['e2e']['ms_per_step'], r['frac'], r['by_kind_ms'])"
timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
) > gpurun_out/c58.log 2>&1
cat gpurun_out/c58.log
