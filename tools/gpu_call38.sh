#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( for s in dep3 l4 l1; do timeout 200 python tools/bench_fprop.py $s 2>&1 | head -8; done
timeout 200 python tools/bench_fprop.py l1 bn 2>&1 | head -8 ) > gpurun_out/c38_fprop.log 2>&1
cat gpurun_out/c38_fprop.log
