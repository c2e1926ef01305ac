#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multistage_gpu.py tests/test_tuned_tiles_gpu.py tests/test_variants_gpu.py -q -p no:cacheprovider 2>&1 | tail -2
timeout 400 python bench.py --arch multistage --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/f1_per_launch_multistage.txt > gpurun_out/f1_bench_multistage.json 2> gpurun_out/f1_bench_ms.err
python -c "
import json; d=json.loads(open('gpurun_out/f1_bench_multistage.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_serial_sum'])"
