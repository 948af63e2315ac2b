#!/bin/bash
OUT=gpurun_out
BARGS="--steps 1 --warmup 3 --no-cpu-baseline --no-fp32 --no-config5"
python bench.py $BARGS > $OUT/plain_decode_r02p.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"nms_topk_stream|refine_stream" -s 6 -c 2 -f -o $OUT/decode_r02p python bench.py $BARGS > $OUT/ncu_decode_r02p.log 2>&1
echo "ncu rc=$?"; ls -la $OUT/decode_r02p.ncu-rep
