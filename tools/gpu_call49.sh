#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( timeout 200 python tools/bench_fprop.py stem 2>&1 | head -16 ) > gpurun_out/c49.log 2>&1
cat gpurun_out/c49.log
