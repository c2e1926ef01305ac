#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --arch multistage --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/n2_bench_ms.json 2> gpurun_out/n2_bench_ms.err
python - <<'PY'
import json
for f in ("n2_bench","n2_bench_ms"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"])
    except Exception as e: print(f,"ERR",e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
