#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2; do
  timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['by_kind_ms']; print('latefusion', d['ms_per_step'], d['e2e']['ms_per_step'], r['pack_weights'])"
done
timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
for b in 1 16; do timeout 100 python tools/eval_latency.py $b 2>&1 | tail -1; done
