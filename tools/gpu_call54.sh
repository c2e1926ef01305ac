#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_tuned_tiles_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
RD_WG_SW128=1 timeout 300 python tools/check_wgrad_sw128.py l1 l2 l3 s3 up2 2>&1 | tail -9
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/c54_per_launch.txt 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(d['ms_per_step'], d['e2e']['ms_per_step'], r['frac'], r['frac_serial_sum'], r['by_kind_ms'])"
timeout 200 python bench.py --arch multistage --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-timing 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
) > gpurun_out/c54.log 2>&1
cat gpurun_out/c54.log
