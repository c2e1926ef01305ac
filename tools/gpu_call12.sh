#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
cp radar_depth_b200/tuned_tiles.json gpurun_out/tuned_tiles_before.json
( time timeout 1200 python tools/autotune.py 16 ) > gpurun_out/c12_autotune_b16.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
( time timeout 900 python tools/autotune.py 8 ) > gpurun_out/c12_autotune_b8.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
( time timeout 900 python tools/autotune.py 8 352 1216 5 ) > gpurun_out/c12_autotune_b8_c5.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
timeout 200 python bench.py --arch multistage --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c12_bench_ms.json 2> gpurun_out/c12_bench_ms.err
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/c12_bench.json gpurun_out/c12_bench_ms.json
grep -h "real\|entries\|REJECTED" gpurun_out/c12_autotune_*.log | head -20
