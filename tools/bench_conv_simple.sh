#!/bin/bash
# one conv shape for ncu: N H W C (3x3 s1, residual + ReLU), 3 launches
python tools/bench_conv.py 0 $1 $2 $3 $4 $4 3 1 3
