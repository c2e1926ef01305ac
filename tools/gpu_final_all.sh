#!/bin/bash
bash tools/gpu_final1.sh
bash tools/gpu_final2.sh
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
