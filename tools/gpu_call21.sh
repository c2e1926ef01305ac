#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_tuned_tiles_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 200 python tools/bench_fprop.py l1 2>&1 | head -3 | cut -c1-220
timeout 200 python tools/bench_fprop.py l2 2>&1 | head -3 | cut -c1-220
timeout 200 python tools/bench_fprop.py stem 2>&1 | head -3 | cut -c1-220
timeout 200 python tools/bench_fprop.py up4 2>&1 | head -3 | cut -c1-220
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c21_bench.json 2> gpurun_out/c21_bench.err
timeout 200 python bench.py --arch multistage --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c21_bench_ms.json 2> gpurun_out/c21_bench_ms.err
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/c21_bench.json gpurun_out/c21_bench_ms.json
