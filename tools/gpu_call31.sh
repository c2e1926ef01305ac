#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for m in 0 1 2; do RD_WG_SW128=$m timeout 300 python tools/check_wgrad_sw128.py 2>&1 | tail -14; done > gpurun_out/c31_sw128.log
cat gpurun_out/c31_sw128.log
