#!/bin/bash
# what the driver runs at round end, in short: GPU suite (-x), smoke(), the default bench line
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/sanity_bench.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/sanity_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
