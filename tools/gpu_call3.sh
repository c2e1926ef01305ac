#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for k in 0 8 12 16 24; do
  RD_DEPTH_SMS=$k timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c3_bench_k$k.json 2> gpurun_out/c3_bench_k$k.err
done
for k in 0 12 24; do
  RD_DEPTH_SMS=$k timeout 200 python bench.py --arch multistage --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/c3_bench_ms_k$k.json 2> gpurun_out/c3_bench_ms_k$k.err
done
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_multistage_gpu.py tests/test_pnp_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/c3_pytest.log
for s in dep1 d16 stem up4 l1; do timeout 120 python tools/bench_fprop.py $s; done > gpurun_out/c3_fprop.txt 2>&1
for s in d16 up4 l1; do timeout 120 python tools/bench_wgrad.py $s; done > gpurun_out/c3_wgrad.txt 2>&1
grep -h ms_per_step gpurun_out/c3_bench_k*.json | python -c "import sys,json; [print(round(json.loads(l)['ms_per_step'],3)) for l in sys.stdin]"
tail -3 gpurun_out/c3_pytest.log
