#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_elementwise_gpu.py -k maxpool -m gpu -q -x -p no:cacheprovider 2>&1 | tail -12
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_multistage_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
for v in 1 0; do echo -n "lf stem_bwd2=$v: "; RD_STEM_BWD2=$v timeout 200 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
for v in 1 0; do echo -n "ms stem_bwd2=$v: "; RD_STEM_BWD2=$v timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['by_kind_ms'])"
