#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_hhrnet_gpu.py -m gpu -x -q -k "chain" > $OUT/chain_tests.log 2>&1; echo "chain tests rc=$?"; tail -3 $OUT/chain_tests.log
for cfg in "X=0" "BRTPE_CHAIN_PF=0" "BRTPE_CHAIN_PF=1" "BRTPE_CHAIN_PF=4" "BRTPE_CHAIN_DEFER=0" "BRTPE_HALO_RES_PREFETCH=1" "BRTPE_CHAIN_G=8" "BRTPE_CHAIN_G=2"; do
  echo "== $cfg"
  env $cfg timeout 60 python tools/bench_chain.py 64 160 160 48 2>&1 | grep "^N="
done
timeout 60 python tools/bench_chain.py 64 320 320 48 10 2>&1 | grep "^N="
timeout 60 python tools/bench_chain.py 16 320 320 48 10 2>&1 | grep "^N="
