#!/bin/bash
# usage: tools/gpu.sh <timeout_s> '<command>'  -- rebuilds the library, then runs the command on a B200 box via gpurun
set -e
cd "$(dirname "$0")/.."
python -m radar_depth_b200.build
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "export PYTHONPATH=.; mkdir -p gpurun_out; $2"
