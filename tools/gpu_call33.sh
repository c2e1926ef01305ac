#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( for m in 0 1; do RD_WG_SW128=$m timeout 300 python tools/check_wgrad_sw128.py 2>&1 | tail -12; done
for m in 1; do for s in l1 l2 d16 stem; do echo "=== mode $m shape $s"; RD_WG_SW128=$m timeout 300 python tools/bench_wgrad.py $s 2>&1 | tail -7; done; done ) > gpurun_out/c33_sw128.log
cat gpurun_out/c33_sw128.log
