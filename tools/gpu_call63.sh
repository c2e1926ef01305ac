#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
cp ab/lib_new.so radar_depth_b200/libradar_depth_b200.so
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_tuned_tiles_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2; do
  for t in old new; do
    cp ab/lib_$t.so radar_depth_b200/libradar_depth_b200.so
    echo -n "$t latefusion: "; timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
  done
done
for t in old new; do
    cp ab/lib_$t.so radar_depth_b200/libradar_depth_b200.so
    echo -n "$t multistage: "; timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
done
