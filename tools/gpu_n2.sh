#!/bin/bash
export NG=${NG:-2}
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/n${NG}_bench.json 2> gpurun_out/n${NG}_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --arch multistage --gpus $NG --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-timing > gpurun_out/n${NG}_bench_ms.json 2> gpurun_out/n${NG}_bench_ms.err
python - <<'PY'
import json
import os
for f in ("n%s_bench" % os.environ.get("NG","2"), "n%s_bench_ms" % os.environ.get("NG","2")):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"])
    except Exception as e: print(f,"ERR",e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
