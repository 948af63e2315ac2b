#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_hhrnet_gpu.py -m gpu -x -q -k "chain" > $OUT/chain_tests.log 2>&1; echo "chain tests rc=$?"; tail -5 $OUT/chain_tests.log
for pdl in 1 0; do
  echo "== PDL=$pdl"
  BRTPE_PDL=$pdl timeout 120 python tools/bench_chain.py 64 160 160 48 2>&1 | grep "^N="
  BRTPE_PDL=$pdl timeout 120 python tools/bench_chain.py 64 320 320 48 10 2>&1 | grep "^N="
done
for g in 1 2 8 16; do
  echo "== G=$g"; BRTPE_CHAIN_G=$g timeout 120 python tools/bench_chain.py 64 160 160 48 2>&1 | grep "CHAIN=2"
done
echo "== no res prefetch"; BRTPE_HALO_RES_PREFETCH=0 timeout 120 python tools/bench_chain.py 64 160 160 48 2>&1 | grep "^N="
echo "== N=8 (L2 resident)"; timeout 120 python tools/bench_chain.py 8 160 160 48 2>&1 | grep "^N="
for cfg in "BRTPE_CHAIN=0 BRTPE_PDL=0" "BRTPE_CHAIN=1 BRTPE_PDL=0" "BRTPE_CHAIN=0 BRTPE_PDL=1" "BRTPE_CHAIN=1 BRTPE_PDL=1" "BRTPE_CHAIN=0 BRTPE_PDL=0"; do
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp32 --no-config5 > $OUT/chain_bench.json 2> $OUT/chain_bench.err
  python - <<P
import json
d=json.loads(open("$OUT/chain_bench.json").read().strip().splitlines()[-1])
print("$cfg value %.1f e2e %.1f ms %.3f halo_frac %.3f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["launches"], d["clocks"]["sm_mhz"]))
P
done
