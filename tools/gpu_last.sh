#!/bin/bash
bash tools/gpu_full.sh r02q
ROUNDS=3 ITERS=60 IDLE=2 timeout 900 python tools/stress_fp32.py 2>&1 | tail -4 | cut -c1-200
