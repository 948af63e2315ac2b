#!/bin/bash
# Standard GPU pass (run under gpurun): parity tests, bench line, per-op / per-phase timing,
# the ncu launch list of the bench command and one `--set full` capture of the dominant conv.
# usage: tools/gpu_round.sh TAG [skip_ncu]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/gpu_tests_$TAG.log
tail -3 $OUT/gpu_tests_$TAG.log
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"
CHUNK=16 python tools/perf_forward.py > $OUT/perf_$TAG.log 2>&1
python tools/perf_step.py > $OUT/perf_step_$TAG.log 2>&1
tail -12 $OUT/perf_step_$TAG.log
if [ -z "$2" ]; then
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 4300 -c 1400 --csv \
      --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
      > $OUT/ncu_launches_$TAG.log 2>&1
  python tools/bench_conv.py 0 16 80 80 96 96 3 1 3 > $OUT/plain_conv_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 \
      -f -o $OUT/halo96_$TAG python tools/bench_conv.py 0 16 80 80 96 96 3 1 3 > $OUT/ncu_halo96_$TAG.log 2>&1
fi
ls $OUT | grep $TAG
