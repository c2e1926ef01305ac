#!/bin/bash
bash tools/gpu_final1.sh
bash tools/gpu_final2.sh
