#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 300 python tests/tools/debug_infer.py > gpurun_out/c17_debug.log 2>&1
cat gpurun_out/c17_debug.log | tail -40
