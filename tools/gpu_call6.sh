#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_variants_gpu.py "tests/test_model_gpu.py::test_segmented_backward_equals_the_single_program" -m gpu -q -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/c6_pytest.log
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c6_bench_n1.json 2> gpurun_out/c6_bench_n1.err
for ov in 1 0; do
  RD_DDP_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-kernel-timing > gpurun_out/c6_bench_n2_ov$ov.json 2> gpurun_out/c6_bench_n2_ov$ov.err
done
for ov in 1 0; do
  RD_DDP_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --arch multistage --gpus 2 --steps 30 --warmup 5 --no-kernel-timing > gpurun_out/c6_bench_ms_n2_ov$ov.json 2> gpurun_out/c6_bench_ms_n2_ov$ov.err
done
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/c6_bench_n1.json gpurun_out/c6_bench_n2_ov1.json gpurun_out/c6_bench_n2_ov0.json gpurun_out/c6_bench_ms_n2_ov1.json gpurun_out/c6_bench_ms_n2_ov0.json
tail -4 gpurun_out/c6_pytest.log
