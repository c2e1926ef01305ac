#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_loop_gpu.py tests/test_model_gpu.py::test_b16_full_size_bf16_tuned_tiles_against_cost_model_tiles_and_oracle tests/test_multistage_gpu.py::test_multistage_b8_bf16_as_benchmarked_against_its_own_fp32_mode -m gpu -q -s -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/c2_pytest.log
for s in dep1 d16 stem up4 l1 dep3; do timeout 120 python tools/bench_fprop.py $s; done > gpurun_out/c2_fprop.txt 2>&1
for s in d16 up4 l1; do timeout 120 python tools/bench_wgrad.py $s; done > gpurun_out/c2_wgrad.txt 2>&1
tail -3 gpurun_out/c2_pytest.log
