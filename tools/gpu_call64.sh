#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
cp ab/lib_new.so radar_depth_b200/libradar_depth_b200.so
timeout 900 python -m pytest tests/test_elementwise_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2; do
  for t in old new; do
    cp ab/lib_$t.so radar_depth_b200/libradar_depth_b200.so
    echo -n "$t latefusion: "; timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['by_kind_ms']; print(d['ms_per_step'], r['join_bwd'], r['head_conv_bwd'])"
  done
done
