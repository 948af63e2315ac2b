#!/bin/bash
OUT=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > $OUT/bench_r02s_n4.json 2> $OUT/bench_r02s_n4.err; echo "rc=$?"
cut -c1-200 $OUT/bench_r02s_n4.json
