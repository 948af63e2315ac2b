#!/bin/bash
OUT=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/gpu_tests_r02r.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/gpu_tests_r02r.log
ROUNDS=3 ITERS=60 IDLE=2 timeout 900 python tools/stress_fp32.py 2>&1 | tail -4 | cut -c1-200
for cfg in "BRTPE_HALO_CG_NARROW=0" "X=0" "BRTPE_HALO_CG_NARROW=0" "X=0"; do
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config5 > $OUT/cg3_bench.json 2> $OUT/cg3_bench.err
  python - <<P
import json
d=json.loads(open("$OUT/cg3_bench.json").read().strip().splitlines()[-1])
print("$cfg value %.1f e2e %.1f ms %.3f halo_frac %.3f fp32 %.1f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["fp32"]["value"], d["clocks"]["sm_mhz"]))
P
done
