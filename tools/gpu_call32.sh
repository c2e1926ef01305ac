#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for m in 0 1; do for s in l1 l2; do echo "=== mode $m shape $s"; RD_WG_SW128=$m timeout 300 python tools/bench_wgrad.py $s 2>&1 | tail -8; done; done > gpurun_out/c32_sw128_bench.log
cat gpurun_out/c32_sw128_bench.log
