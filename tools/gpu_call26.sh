#!/bin/bash
export PYTHONPATH=$PWD
for i in 1 2; do for gc in 1 0; do echo -n "ms gc=$gc: "; RD_WGRAD_GCOPY=$gc timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done; done
for gc in 1 0; do echo -n "lf gc=$gc: "; RD_WGRAD_GCOPY=$gc timeout 200 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
