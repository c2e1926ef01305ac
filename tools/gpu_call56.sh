#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_infer_pack_hash_gpu.py tests/test_model_gpu.py tests/test_pnp_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
for b in 1 16; do timeout 100 python tools/eval_latency.py $b 2>&1 | tail -1; RD_INFER_PACK_HASH=0 timeout 100 python tools/eval_latency.py $b 2>&1 | tail -1; done
) > gpurun_out/c56.log 2>&1
cat gpurun_out/c56.log
