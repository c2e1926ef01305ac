#!/bin/bash
# ncu evidence of round 2 (final code): launch list of ~2 training steps + --set full of selected launches
export PYTHONPATH=$PWD
mkdir -p gpurun_out
RADAR_DEPTH_B200_GRAPHS=0 timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 800 -c 520 --csv --log-file gpurun_out/f2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-timing > gpurun_out/f2_launches_bench.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/f2_full python tools/run_launch.py 1 16 conv_f:layer1.0.conv1 conv_f:layer1.0.conv2 conv_f:layer3.1.conv1 conv_d:layer1.0.conv2 wgrad:layer1.0.conv1 wgrad:layer2.0.conv2 wgrad:layer2.0.conv1 wgrad:layer4.1.conv1 conv_f:stem maxpool maxpool_bwd_apply join_bwd:layer1.0 bn_bwd_apply:u1 > gpurun_out/f2_full.log 2>&1
tail -3 gpurun_out/f2_full.log
ls -la gpurun_out/f2_full.ncu-rep gpurun_out/f2_launches.csv
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/f2_bench_e2e.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/f2_bench_e2e.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e'])"
