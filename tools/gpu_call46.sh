#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -5
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/c46_per_launch.txt 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['by_kind_ms'])"
timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
for b in 1 16; do for d in 0 20; do echo -n "depth_sms=$d "; RD_DEPTH_SMS=$d timeout 100 python tools/eval_latency.py $b 2>&1 | tail -1; done; done
) > gpurun_out/c46.log 2>&1
cat gpurun_out/c46.log
