#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_variants_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/c5_pytest.log
for k in 0 24; do
  RD_DEPTH_SMS=$k timeout 200 python bench.py --arch multistage --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c5_bench_ms_k$k.json 2> gpurun_out/c5_bench_ms_k$k.err
done
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/c5_bench_ms_k*.json
tail -8 gpurun_out/c5_pytest.log
