#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_elementwise_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/c10_pytest_a.log
tail -3 gpurun_out/c10_pytest_a.log
timeout 120 python tools/launch_gap.py l1 > gpurun_out/c10_gap_l1.log 2>&1
timeout 120 python tools/launch_gap.py dep1 > gpurun_out/c10_gap_dep1.log 2>&1
for pdl in 1 0; do
  RD_PDL=$pdl timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c10_bench_pdl$pdl.json 2> gpurun_out/c10_bench_pdl$pdl.err
  RD_PDL=$pdl timeout 200 python bench.py --arch multistage --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c10_bench_ms_pdl$pdl.json 2> gpurun_out/c10_bench_ms_pdl$pdl.err
done
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/c10_bench_*.json
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/c10_pytest_full.log
tail -4 gpurun_out/c10_pytest_full.log
head -8 gpurun_out/c10_gap_l1.log gpurun_out/c10_gap_dep1.log
