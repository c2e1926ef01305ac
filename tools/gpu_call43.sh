#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_multistage_gpu.py tests/test_train_loop_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -4
for i in 1 2; do timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
) > gpurun_out/c43.log 2>&1
cat gpurun_out/c43.log
