#!/bin/bash
# round 2, GPU call S (2 GPUs): real-NCCL tests and the default bench line at N=2 on the tuned kernels
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -v ) > gpurun_out/r2s_pytest_multigpu.log 2>&1; tail -8 gpurun_out/r2s_pytest_multigpu.log
P=$((20000 + RANDOM % 20000))
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 6 --warmup 3 ) > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.err; tail -c 400 gpurun_out/r2s_bench_n2.err
