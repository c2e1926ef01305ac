#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
cp radar_depth_b200/tuned_tiles.json /tmp/merged.json
for rep in 1 2; do
  for t in old merged; do
    if [ $t = old ]; then cp gpurun_out_in/tuned_tiles_before_c60.json radar_depth_b200/tuned_tiles.json; else cp /tmp/merged.json radar_depth_b200/tuned_tiles.json; fi
    echo -n "$t: "; timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
  done
done
