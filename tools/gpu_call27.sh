#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( time timeout 1500 python tools/autotune.py 16 ) > gpurun_out/c27_autotune_b16.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
( time timeout 1200 python tools/autotune.py 8 ) > gpurun_out/c27_autotune_b8.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
( time timeout 1200 python tools/autotune.py 8 352 1216 5 ) > gpurun_out/c27_autotune_b8_c5.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
grep -h "REJECTED\|NO VALID\|real" gpurun_out/c27_autotune_*.log | head
for i in 1 2; do timeout 200 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
