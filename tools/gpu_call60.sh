#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
cp radar_depth_b200/tuned_tiles.json gpurun_out/tuned_tiles_before_c60.json
timeout 1500 python tools/autotune.py 16 > gpurun_out/c60_autotune_b16.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
timeout 1500 python tools/autotune.py 8 > gpurun_out/c60_autotune_b8.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
timeout 1500 python tools/autotune.py 8 352 1216 5 > gpurun_out/c60_autotune_b8_c5.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
tail -1 gpurun_out/c60_autotune_b16.log
timeout 600 python -m pytest tests/test_tuned_tiles_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -2
for i in 1 2; do timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "--- with the table from before:"
cp gpurun_out/tuned_tiles_before_c60.json radar_depth_b200/tuned_tiles.json
for i in 1 2; do timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
