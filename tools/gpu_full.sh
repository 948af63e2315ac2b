#!/bin/bash
# full GPU pass: every -m gpu test, smoke(), the default bench line
OUT=gpurun_out; TAG=${1:-r02m}
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/gpu_tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/gpu_tests_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 $OUT/bench_$TAG.json
