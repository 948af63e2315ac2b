#!/bin/bash
# Round-2 GPU pass (under gpurun): bench line, then -- only after the plain commands exited 0 -- the ncu
# launch list of the same bench command and one `--set full` capture of the dominant conv launch shape.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
BARGS="--steps 1 --warmup 3 --no-cpu-baseline --no-fp32 --no-config5"
python bench.py $BARGS > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:"conv_|fuse_sum|nhwc|stem_im2col|aggregate|nms_topk|topk_merge|group_ae|adjust_kernel|scores_kernel|refine|prepack" \
    -c 3000 --csv --log-file $OUT/launches_$TAG.csv python bench.py $BARGS > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
python tools/bench_conv.py 0 64 160 160 48 48 3 1 3 > $OUT/plain_conv48_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 \
    -f -o $OUT/halo48_$TAG python tools/bench_conv.py 0 64 160 160 48 48 3 1 3 > $OUT/ncu_halo48_$TAG.log 2>&1
echo "ncu halo48 rc=$?"
python tools/bench_conv.py 0 64 80 80 96 96 3 1 3 > $OUT/plain_conv96_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 \
    -f -o $OUT/halo96_$TAG python tools/bench_conv.py 0 64 80 80 96 96 3 1 3 > $OUT/ncu_halo96_$TAG.log 2>&1
echo "ncu halo96 rc=$?"
ls -la $OUT | grep $TAG
