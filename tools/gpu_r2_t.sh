#!/bin/bash
# round 2, GPU call T (8 GPUs): the default bench line at N=8 on the tuned kernels (primary project-sharded, secondary c4 time-sharded)
mkdir -p gpurun_out
P=$((20000 + RANDOM % 20000))
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 6 --warmup 3 ) > gpurun_out/r2t_bench_n8.json 2> gpurun_out/r2t_bench_n8.err; tail -c 400 gpurun_out/r2t_bench_n8.err
