#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python tools/autotune.py 8 > gpurun_out/c4_autotune_b8.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
timeout 900 python tools/autotune.py 8 352 1216 5 > gpurun_out/c4_autotune_b8_c5.log 2>&1
cp gpurun_out/tuned_tiles.json radar_depth_b200/tuned_tiles.json
for k in 0 24 32 40; do
  RD_DEPTH_SMS=$k timeout 200 python bench.py --arch multistage --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c4_bench_ms_k$k.json 2> gpurun_out/c4_bench_ms_k$k.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/c4_launches_latefusion.txt > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/c4_bench_ms_k*.json gpurun_out/c4_bench.json
tail -2 gpurun_out/c4_autotune_b8.log gpurun_out/c4_autotune_b8_c5.log
