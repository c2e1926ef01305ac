#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 300 python tools/bench_fprop.py l2 raw 4x51 2x126 5x94 3x83 1x126 2x62 > gpurun_out/c18_fprop_l2.log 2>&1
grep -v "ctas " gpurun_out/c18_fprop_l2.log | cut -c1-250
