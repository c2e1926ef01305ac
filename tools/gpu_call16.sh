#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/c16_pytest_full.log
tail -5 gpurun_out/c16_pytest_full.log
