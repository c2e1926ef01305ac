#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
export RD_TUNE_LANES_ONLY=1
timeout 1200 python tools/autotune.py 16 > gpurun_out/c47_autotune_b16.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
timeout 1200 python tools/autotune.py 8 > gpurun_out/c47_autotune_b8.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
timeout 1200 python tools/autotune.py 8 352 1216 5 > gpurun_out/c47_autotune_b8_c5.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
tail -2 gpurun_out/c47_autotune_b16.log | cut -c1-200
timeout 600 python -m pytest tests/test_tuned_tiles_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/c47_per_launch.txt 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(d['ms_per_step'], d['e2e']['ms_per_step'], r['frac'], r['frac_serial_sum'], r['frac_step'], r['lanes'], r['by_kind_ms'])"
timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
