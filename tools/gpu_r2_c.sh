#!/bin/bash
# round 2, GPU call C (2 GPUs): real-NCCL tests, bench at N=2 (primary project-sharded + secondary time-sharded c4)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c_topo.txt 2>&1
( time timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -v ) > gpurun_out/r2c_pytest_multigpu.log 2>&1; tail -8 gpurun_out/r2c_pytest_multigpu.log
P=$((20000 + RANDOM % 20000))
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 8 --warmup 3 ) > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; tail -c 400 gpurun_out/r2c_bench_n2.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) tools/pcie_bw_all.py > gpurun_out/r2c_pcie_n2.json 2> gpurun_out/r2c_pcie_n2.err; cat gpurun_out/r2c_pcie_n2.json
