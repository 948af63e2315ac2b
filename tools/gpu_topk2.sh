#!/bin/bash
OUT=gpurun_out
run() {
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp32 --no-config5 > $OUT/topk2_bench.json 2> $OUT/topk2_bench.err
  python - <<P
import json
d=json.loads(open("$OUT/topk2_bench.json").read().strip().splitlines()[-1])
r=d["roofline_decode"]
print("$1 top_k %.4f ms frac %.3f | value %.1f" % (r["top_k"]["ms"], r["top_k"]["frac"], d["value"]))
P
}
run default
BRTPE_TOPK_HINT=500 run hint500
BRTPE_TOPK_HINT=2000 run hint2000
BRTPE_TOPK_HINT=10000 run hint10000
BRTPE_TOPK_SLEEP=4000 run sleep4000
BRTPE_TOPK_SLEEP=10000 run sleep10000
BRTPE_TOPK_HINT=2000 BRTPE_TOPK_SLEEP=4000 run hint2000_sleep4000
run default
