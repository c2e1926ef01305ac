#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for n in l1 l4 dep1; do timeout 120 python tools/launch_gap.py $n > gpurun_out/c9_gap_$n.log 2>&1; done
timeout 120 python tools/launch_gap.py l1 bn > gpurun_out/c9_gap_l1bn.log 2>&1
cat gpurun_out/c9_gap_*.log
