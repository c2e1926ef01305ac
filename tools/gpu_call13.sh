#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
RADAR_DEPTH_B200_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 800 -c 520 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-timing > gpurun_out/c13_ncu_bench.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02_launches.csv "round 2: latefusion b=16 bf16, ~2 training steps (eager launches), ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none" --traffic-json gpurun_out/r02_conv_traffic.json > gpurun_out/r02_launches_summary.txt 2>&1
RADAR_DEPTH_B200_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1700 -c 1100 --csv --log-file gpurun_out/r02_launches_multistage.csv python bench.py --arch multistage --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-timing > gpurun_out/c13_ncu_bench_ms.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02_launches_multistage.csv "round 2: multistage-fixs b=8 bf16, ~2 training steps (eager launches)" > gpurun_out/r02_launches_multistage_summary.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_full python tools/run_launch.py 1 16 conv_f:layer1.0.conv1 conv_f:layer2.1.conv2 conv_f:layer3.1.conv1 conv_d:layer1.0.conv2 wgrad:layer1.0.conv1 wgrad:layer2.1.conv2 wgrad:stem conv_f:stem wgrad:decoder.layer4.upper_branch.conv2 maxpool_bwd bn_bwd_apply:stem join_bwd:layer1.0 > gpurun_out/c13_ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_full.ncu-rep "r02" > gpurun_out/r02_ncu_full.txt 2>&1
ls -la gpurun_out/r02_full.ncu-rep; rm -f gpurun_out/r02_full.ncu-rep
head -30 gpurun_out/r02_launches_summary.txt
