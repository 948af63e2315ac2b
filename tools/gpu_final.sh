#!/bin/bash
# final pass of a round: every GPU test, smoke(), then tools/gpu_round2.sh (bench line, ncu launch list, two
# --set full captures)
TAG=${1:-r02p}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=serial --format=csv,noheader
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/gpu_tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/gpu_tests_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
bash tools/gpu_round2.sh $TAG
