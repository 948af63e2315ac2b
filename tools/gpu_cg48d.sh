#!/bin/bash
for cfg in "X=0" "BRTPE_HALO_TMA_OUT=1" "BRTPE_HALO_RES_PREFETCH=0" "BRTPE_HALO_A_STAGES=2" "BRTPE_HALO_A_STAGES=4" "BRTPE_HALO_RES_STAGED=1" "X=1"; do
  echo "== $cfg"; env $cfg python tools/halo_prof.py 64 160 160 48 2>&1 | grep -E "^N=|mma.wait_acc|epi.wait"
done
for cfg in "X=0" "BRTPE_HALO_TMA_OUT=1" "BRTPE_HALO_A_STAGES=4"; do
echo "== 256->48 $cfg"; env $cfg python tools/bench_conv.py 0 64 160 160 256 48 3 1 2>&1 | tail -1
echo "== 64->64 $cfg"; env $cfg python tools/bench_conv.py 0 64 160 160 64 64 3 1 2>&1 | tail -1
done
