#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dataset_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/c14_pytest.log
cat gpurun_out/c14_pytest.log
