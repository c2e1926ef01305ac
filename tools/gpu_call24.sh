#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_tuned_tiles_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15
for n in l1 d16; do timeout 120 python tools/bench_wgrad.py $n 2>&1 | tail -4 | cut -c1-200; done
for gc in 1 0; do RD_WGRAD_GCOPY=$gc timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
