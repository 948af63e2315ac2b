#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_golden_gpu.py tests/test_configs_gpu.py -m gpu -x -q > $OUT/refine_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/refine_tests.log
BASE=$PWD/tools/ab/libbrtpe_base.so
run() {
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp32 > $OUT/refine_bench.json 2> $OUT/refine_bench.err
  python - <<P
import json
d=json.loads(open("$OUT/refine_bench.json").read().strip().splitlines()[-1])
r=d["roofline_decode"]; c=d["config5_decode"]
print("$1 dense refine %.4f ms frac %.3f | config5 top_k %.3f refine %.3f match %.3f total %.3f img/s %.0f parse_frac %.3f | value %.1f" % (r["refine"]["ms"], r["refine"]["frac"], c["ms"]["top_k"], c["ms"]["refine"], c["ms"]["match"], c["ms"]["parse_total"], c["value"], c["parse_frac"], d["value"]))
P
}
BRTPE_LIB=$BASE run base
run new
BRTPE_LIB=$BASE run base
run new
