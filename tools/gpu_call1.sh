#!/bin/bash
# round-2 GPU call 1: full GPU suite (no -x), benches (latefusion / multistage / fp32), cuDNN bar
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/c1_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --dump-launches gpurun_out/c1_launches_latefusion.txt > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
timeout 300 python bench.py --arch multistage --steps 20 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/c1_launches_multistage.txt > gpurun_out/c1_bench_ms.json 2> gpurun_out/c1_bench_ms.err
timeout 300 python bench.py --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_fp32.json 2> gpurun_out/c1_bench_fp32.err
timeout 300 python tests/tools/cudnn_bar.py --arch latefusion --steps 30 --out gpurun_out/c1_cudnn_bar_latefusion.json > gpurun_out/c1_cudnn.log 2>&1
timeout 300 python tests/tools/cudnn_bar.py --arch multistage --steps 30 --out gpurun_out/c1_cudnn_bar_multistage.json >> gpurun_out/c1_cudnn.log 2>&1
tail -5 gpurun_out/c1_pytest.log
