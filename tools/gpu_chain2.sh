#!/bin/bash
for cfg in "X=0" "BRTPE_HALO_A_STAGES=2" "BRTPE_CHAIN_DBG=1" "BRTPE_CHAIN_DBG=2" "BRTPE_CHAIN_DBG=3" "BRTPE_CHAIN_DBG=3 BRTPE_HALO_RES_PREFETCH=0"; do
  echo "== $cfg"
  env $cfg timeout 120 python tools/bench_chain.py 64 160 160 48 2>&1 | grep "^N="
  env $cfg timeout 120 python tools/bench_chain.py 8 160 160 48 2>&1 | grep "^N="
done
