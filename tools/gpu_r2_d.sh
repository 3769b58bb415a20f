#!/bin/bash
# round 2, GPU call D (8 GPUs): bench at N=8 with the library's own job timeline of rank 0, aggregate PCIe ceiling
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2d_topo.txt 2>&1; nproc >> gpurun_out/r2d_topo.txt; numactl -H >> gpurun_out/r2d_topo.txt 2>&1
P=$((20000 + RANDOM % 20000))
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P tools/pcie_bw_all.py > gpurun_out/r2d_pcie_n8.json 2> gpurun_out/r2d_pcie_n8.err; cat gpurun_out/r2d_pcie_n8.json
( time timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 8 --steps 8 --warmup 3 ) > gpurun_out/r2d_bench_n8.json 2> gpurun_out/r2d_bench_n8.err; tail -c 400 gpurun_out/r2d_bench_n8.err
TDG_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus 8 --steps 2 --warmup 1 --no-secondary > gpurun_out/r2d_bench_n8_trace.json 2> gpurun_out/r2d_bench_n8_trace.err
grep -E "timeline job|plan lane" gpurun_out/r2d_bench_n8_trace.err | tail -120 > gpurun_out/r2d_n8_timeline.log; rm -f gpurun_out/r2d_bench_n8_trace.err
