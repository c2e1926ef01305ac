#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( for d in 12 16 18 20; do echo -n "depth_sms=$d latefusion: "; RD_DEPTH_SMS=$d timeout 200 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
for d in 16 18 20; do echo -n "depth_sms=$d multistage: "; RD_DEPTH_SMS=$d timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
) > gpurun_out/c48.log 2>&1
cat gpurun_out/c48.log
