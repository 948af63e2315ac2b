#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_hhrnet_gpu.py tests/test_fullsize_gpu.py tests/test_student_gpu.py tests/test_golden_gpu.py -m gpu -x -q > $OUT/epi8b_tests.log 2>&1; echo "tests rc=$?"; tail -2 $OUT/epi8b_tests.log
for sh in "0 64 160 160 64 256 1 1" "0 64 160 160 256 64 1 1" "0 64 320 320 32 64 1 1"; do
  for cfg in "BRTPE_UMMA_EPI8=0" "BRTPE_UMMA_EPI8=1"; do echo -n "$cfg: "; env $cfg timeout 60 python tools/bench_conv.py $sh 2>&1 | tail -1; done
done
for cfg in "BRTPE_UMMA_EPI8=0" "BRTPE_UMMA_EPI8=1" "BRTPE_UMMA_EPI8=0" "BRTPE_UMMA_EPI8=1"; do
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config5 > $OUT/epi8b_bench.json 2> $OUT/epi8b_bench.err
  python - <<P
import json
d=json.loads(open("$OUT/epi8b_bench.json").read().strip().splitlines()[-1])
print("$cfg value %.1f e2e %.1f ms %.3f other_frac %.3f fp32 %.1f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline_other_convs"]["frac"], d["fp32"]["value"], d["clocks"]["sm_mhz"]))
P
done
