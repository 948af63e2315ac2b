#!/bin/bash
OUT=gpurun_out
echo skip tests
BASE=$PWD/tools/ab/libbrtpe_base.so
for sh in "0 64 160 160 64 256 1 1" "0 64 160 160 256 64 1 1" "0 64 320 320 48 32 1 1" "0 64 80 80 96 48 1 1" "0 64 160 160 48 96 3 2" "0 64 20 20 384 384 3 1" "0 64 160 160 48 48 3 1"; do
  echo "== $sh"
  echo -n "base: "; BRTPE_LIB=$BASE timeout 60 python tools/bench_conv.py $sh 2>&1 | tail -1
  echo -n "new : "; timeout 60 python tools/bench_conv.py $sh 2>&1 | tail -1
done
for lib in base new base new; do
  if [ $lib = base ]; then export BRTPE_LIB=$BASE; else unset BRTPE_LIB; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp32 --no-config5 > $OUT/epi8_bench.json 2> $OUT/epi8_bench.err
  python - <<P
import json
d=json.loads(open("$OUT/epi8_bench.json").read().strip().splitlines()[-1])
print("$lib value %.1f e2e %.1f ms %.3f halo_frac %.3f other_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline_other_convs"]["frac"], d["clocks"]["sm_mhz"]))
P
done
