#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
cp radar_depth_b200/tuned_tiles.json gpurun_out/tuned_tiles_before.json
( time timeout 1200 python tools/autotune.py 16 ) > gpurun_out/c25_autotune_b16.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
( time timeout 900 python tools/autotune.py 8 ) > gpurun_out/c25_autotune_b8.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
( time timeout 900 python tools/autotune.py 8 352 1216 5 ) > gpurun_out/c25_autotune_b8_c5.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
grep -h "REJECTED\|NO VALID" gpurun_out/c25_autotune_*.log | head
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err
timeout 200 python bench.py --arch multistage --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c25_bench_ms.json 2> gpurun_out/c25_bench_ms.err
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/c25_bench.json gpurun_out/c25_bench_ms.json
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/c25_pytest_full.log
tail -3 gpurun_out/c25_pytest_full.log
grep -h " w|" gpurun_out/c25_autotune_b16.log | cut -c1-200 | head -50
