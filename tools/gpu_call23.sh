#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/c23_pytest_full.log
tail -3 gpurun_out/c23_pytest_full.log
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err
timeout 200 python bench.py --arch multistage --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c23_bench_ms.json 2> gpurun_out/c23_bench_ms.err
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/c23_bench.json gpurun_out/c23_bench_ms.json
