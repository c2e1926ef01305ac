#!/bin/bash
# final single-GPU evidence run of round 2: tests, bench lines (both architectures, parity mode, reference arm), latency
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/f1_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --dump-launches gpurun_out/f1_per_launch_latefusion.txt > gpurun_out/f1_bench_n1.json 2> gpurun_out/f1_bench_n1.err
timeout 400 python bench.py --arch multistage --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/f1_per_launch_multistage.txt > gpurun_out/f1_bench_multistage.json 2> gpurun_out/f1_bench_ms.err
timeout 400 python bench.py --precision fp32 --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f1_bench_fp32.json 2> gpurun_out/f1_bench_fp32.err
timeout 400 python bench.py --impl reference --gpus 1 --steps 4 --warmup 1 > gpurun_out/f1_bench_reference.json 2> gpurun_out/f1_bench_ref.err
( timeout 100 python tools/eval_latency.py 1 | tail -1; timeout 100 python tools/eval_latency.py 16 | tail -1 ) > gpurun_out/f1_eval_latency.txt 2>&1
( for m in 0 1; do RD_WG_SW128=$m timeout 300 python tools/check_wgrad_sw128.py 2>&1 | tail -20; done ) > gpurun_out/f1_wgrad_sw128_check.txt 2>&1
cat gpurun_out/f1_pytest.log
python - <<'PY'
import json
for f in ("f1_bench_n1","f1_bench_multistage","f1_bench_fp32","f1_bench_reference"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(f, d.get("ms_per_step"), d.get("value"), d.get("e2e",{}).get("value"), r.get("frac"), r.get("frac_serial_sum"), d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
cat gpurun_out/f1_eval_latency.txt
