#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -k "inference_program" -m gpu -q -x -rP -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/c15_pytest.log
grep -h "bf16 inference\|passed\|failed" gpurun_out/c15_pytest.log
for b in 1 16; do for f in 1 0; do RD_INFER_FOLD=$f timeout 200 python tools/eval_latency.py $b > gpurun_out/c15_eval_b${b}_fold$f.log 2>&1; tail -1 gpurun_out/c15_eval_b${b}_fold$f.log; done; done
timeout 900 python -m pytest tests/test_variants_gpu.py tests/test_multistage_gpu.py tests/test_pnp_gpu.py tests/test_train_loop_gpu.py tests/test_metrics_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
