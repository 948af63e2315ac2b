#!/bin/bash
# Ablation of conv_halo_kernel (BRTPE_HALO_DBG bits: 1 no MMA, 2 no epilogue, 4 no weight loads,
# 8 no activation loads): which resource bounds each layer shape.
for shape in "16 160 160 48 48" "16 80 80 96 96" "16 40 40 192 192" "16 20 20 384 384" "32 80 80 96 96"; do
  for dbg in 0 1 2 3 4 6 12 14 13; do
    echo -n "dbg=$dbg  "
    BRTPE_HALO_DBG=$dbg python tools/bench_conv.py 0 $shape 3 1 20 2>&1 | tail -1
  done
done
