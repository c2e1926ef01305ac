#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fprop -s 2 -c 1 -f -o gpurun_out/c50_stem_fprop python tools/bench_fprop.py stem > gpurun_out/c50_ncu.log 2>&1
tail -3 gpurun_out/c50_ncu.log
ls -la gpurun_out/c50_stem_fprop.ncu-rep
