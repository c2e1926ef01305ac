#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for n in l1 l2 l4 up4 d16 stem; do timeout 120 python tools/bench_wgrad.py $n > gpurun_out/c11_wgrad_$n.log 2>&1; done
grep -h "timeline\|TFLOP\|neither\|Error\|error" gpurun_out/c11_wgrad_*.log | cut -c1-220
