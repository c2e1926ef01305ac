#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 python bench.py --dump-launches gpurun_out/r02_per_launch_latefusion_final.txt > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/c22_bench.err
timeout 600 python bench.py --arch multistage --dump-launches gpurun_out/r02_per_launch_multistage_final.txt > gpurun_out/r02_bench_multistage_b8_final.json 2> gpurun_out/c22_bench_ms.err
timeout 600 python bench.py --precision fp32 --no-cpu-baseline > gpurun_out/r02_bench_fp32_parity_mode_final.json 2> gpurun_out/c22_bench_fp32.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm_final.json 2> gpurun_out/c22_bench_ref.err
for b in 1 16; do timeout 200 python tools/eval_latency.py $b >> gpurun_out/r02_eval_latency_final.txt 2>&1; done
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_n1_final.json gpurun_out/r02_bench_multistage_b8_final.json gpurun_out/r02_bench_fp32_parity_mode_final.json
cat gpurun_out/r02_eval_latency_final.txt
