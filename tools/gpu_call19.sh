#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_tuned_tiles_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 200 python tools/bench_fprop.py l1 2>&1 | head -8 | cut -c1-220
timeout 200 python tools/bench_fprop.py l2 raw 4x51 2x126 1x126 2>&1 | grep -v "ctas " | cut -c1-220
timeout 200 python tools/bench_fprop.py l4 2>&1 | head -8 | cut -c1-220
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c19_bench_pre.json 2> gpurun_out/c19_bench.err
cp radar_depth_b200/tuned_tiles.json gpurun_out/tuned_tiles_before.json
( time timeout 1200 python tools/autotune.py 16 ) > gpurun_out/c19_autotune_b16.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
( time timeout 900 python tools/autotune.py 8 ) > gpurun_out/c19_autotune_b8.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
( time timeout 900 python tools/autotune.py 8 352 1216 5 ) > gpurun_out/c19_autotune_b8_c5.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c19_bench.json 2>> gpurun_out/c19_bench.err
timeout 200 python bench.py --arch multistage --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c19_bench_ms.json 2> gpurun_out/c19_bench_ms.err
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/c19_bench_pre.json gpurun_out/c19_bench.json gpurun_out/c19_bench_ms.json
