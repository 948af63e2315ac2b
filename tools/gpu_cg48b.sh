#!/bin/bash
OUT=gpurun_out
BRTPE_HALO_CG=2 timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -x -q > $OUT/cg2_tests.log 2>&1; echo "conv tests with CG=2 rc=$?"; tail -3 $OUT/cg2_tests.log
for sh in "0 64 320 320 48 48 3 1 10" "0 64 160 160 64 64 3 1" "0 64 160 160 256 48 3 1" "0 64 160 160 48 96 3 2" "0 64 160 160 48 48 3 2" "0 64 80 80 48 48 3 1" "0 8 160 160 48 48 3 1" "0 2 160 160 48 48 3 1" "0 64 320 320 64 64 3 2 10"; do
  echo "== $sh"
  echo -n "cg default: "; timeout 60 python tools/bench_conv.py $sh 2>&1 | tail -1
  echo -n "cg2       : "; BRTPE_HALO_CG=2 timeout 60 python tools/bench_conv.py $sh 2>&1 | tail -1
done
for cfg in "X=0" "BRTPE_HALO_CG=2" "X=0" "BRTPE_HALO_CG=2"; do
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp32 --no-config5 > $OUT/cg2_bench.json 2> $OUT/cg2_bench.err
  python - <<P
import json
d=json.loads(open("$OUT/cg2_bench.json").read().strip().splitlines()[-1])
print("$cfg value %.1f e2e %.1f ms %.3f halo_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
P
done
