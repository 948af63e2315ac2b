#!/bin/bash
OUT=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_r02p_n2.json 2> $OUT/bench_r02p_n2.err; echo "rc=$?"
cut -c1-300 $OUT/bench_r02p_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $OUT/bench_r02p_ref_n2.json 2> $OUT/bench_r02p_ref_n2.err; echo "ref rc=$?"
cut -c1-300 $OUT/bench_r02p_ref_n2.json
