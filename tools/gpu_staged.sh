#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_hhrnet_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > $OUT/staged_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/staged_tests.log
for sh in "0 64 160 160 48 48 3 1" "0 64 320 320 48 48 3 1 10" "0 64 160 160 64 64 3 1" "0 8 160 160 48 48 3 1"; do
  for cfg in "BRTPE_HALO_RES_STAGED=0" "BRTPE_HALO_RES_STAGED=1" "BRTPE_HALO_RES_STAGED=1 BRTPE_HALO_RES_PREFETCH=0" "BRTPE_HALO_RES_STAGED=0"; do
    echo -n "$cfg: "; env $cfg timeout 60 python tools/bench_conv.py $sh 2>&1 | tail -1
  done
done
for cfg in "BRTPE_HALO_RES_STAGED=0" "BRTPE_HALO_RES_STAGED=1" "BRTPE_HALO_RES_STAGED=0" "BRTPE_HALO_RES_STAGED=1"; do
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp32 --no-config5 > $OUT/staged_bench.json 2> $OUT/staged_bench.err
  python - <<P
import json
d=json.loads(open("$OUT/staged_bench.json").read().strip().splitlines()[-1])
print("$cfg value %.1f e2e %.1f ms %.3f halo_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
P
done
