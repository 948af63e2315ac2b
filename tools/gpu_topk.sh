#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_golden_gpu.py -m gpu -x -q > $OUT/topk_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/topk_tests.log
BASE=$PWD/tools/ab/libbrtpe_base.so
run() {
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp32 > $OUT/topk_bench.json 2> $OUT/topk_bench.err
  python - <<P
import json
d=json.loads(open("$OUT/topk_bench.json").read().strip().splitlines()[-1])
r=d["roofline_decode"]; c=d["config5_decode"]
print("$1 top_k %.4f ms frac %.3f refine %.4f agg %.4f | config5 top_k %.3f refine %.3f match %.3f total %.3f img/s %.0f | value %.1f" % (r["top_k"]["ms"], r["top_k"]["frac"], r["refine"]["ms"], r["aggregate"]["ms"], c["ms"]["top_k"], c["ms"]["refine"], c["ms"]["match"], c["ms"]["parse_total"], c["value"], d["value"]))
P
}
BRTPE_LIB=$BASE run base
run new
BRTPE_TOPK_SR=8 BRTPE_TOPK_NS=4 run new_sr8_ns4
BRTPE_TOPK_SR=8 BRTPE_TOPK_NS=8 run new_sr8_ns8
BRTPE_TOPK_NC=4 run new_nc4
BRTPE_LIB=$BASE run base
