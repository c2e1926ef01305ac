#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_tuned_tiles_gpu.py "tests/test_model_gpu.py::test_segmented_backward_equals_the_single_program" "tests/test_model_gpu.py::test_train_step_parity_fp32_mode" -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/c8_pytest.log
for n in l1 l2 l4 up4; do timeout 120 python tools/bench_fprop.py $n 2>&1 | head -8 > gpurun_out/c8_fprop_$n.log; done
timeout 120 python tools/bench_fprop.py l1 bn 2>&1 | head -8 > gpurun_out/c8_fprop_l1bn.log
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err
tail -3 gpurun_out/c8_pytest.log; grep -o '"ms_per_step": [0-9.]*' gpurun_out/c8_bench.json | head -2
