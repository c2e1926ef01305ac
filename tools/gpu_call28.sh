#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_tuned_tiles_gpu.py tests/test_variants_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8
for n in up4 up3; do timeout 120 python tools/bench_wgrad.py $n 2>&1 | tail -1 | cut -c1-200; done
( time timeout 1500 python tools/autotune.py 16 ) > gpurun_out/c28_autotune_b16.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
( time timeout 1200 python tools/autotune.py 8 ) > gpurun_out/c28_autotune_b8.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
( time timeout 1200 python tools/autotune.py 8 352 1216 5 ) > gpurun_out/c28_autotune_b8_c5.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
grep -h "REJECTED\|NO VALID" gpurun_out/c28_autotune_*.log | head
grep -h "up5x5 .* w|" gpurun_out/c28_autotune_b16.log | cut -c1-200
timeout 200 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
