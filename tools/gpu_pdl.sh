#!/bin/bash
# PDL A/B on one box: parity tests with PDL on, then the bench with BRTPE_PDL=0 / 1 alternating.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_hhrnet_gpu.py tests/test_fullsize_gpu.py tests/test_golden_gpu.py -m gpu -x -q > $OUT/pdl_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/pdl_tests.log
for v in 0 1 0 1; do
  BRTPE_PDL=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp32 --no-config5 > $OUT/pdl_bench_$v.json 2> $OUT/pdl_bench_$v.err
  python - <<P
import json
d=json.loads(open("$OUT/pdl_bench_$v.json").read().strip().splitlines()[-1])
print("PDL=$v value %.1f e2e %.1f ms %.3f halo_frac %.3f clocks %s parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d.get("parity_checked")))
P
done
