#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/c7_pytest.log
for n in l1 l2 l4 up4 up3 d16; do timeout 120 python tools/bench_wgrad.py $n > gpurun_out/c7_wgrad_$n.log 2>&1; done
for n in l1 l2 l3 l4 stem up4 dep1; do timeout 120 python tools/bench_fprop.py $n > gpurun_out/c7_fprop_$n.log 2>&1; done
timeout 120 python tools/bench_fprop.py l1 bn > gpurun_out/c7_fprop_l1bn.log 2>&1
tail -5 gpurun_out/c7_pytest.log
