#!/bin/bash
# round 2, GPU call A (1 GPU): full -m gpu suite at HEAD, first timings of configs 2/3/4, PCIe ceiling
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
nproc >> gpurun_out/r2a_smi.txt; lscpu | head -20 >> gpurun_out/r2a_smi.txt
( time timeout 900 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2a_pytest_gpu.log
for w in c2 c3 c4; do
  ( time timeout 420 python bench.py --workload $w --steps 2 --warmup 1 ) > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err
  echo "$w rc=$?"; tail -c 400 gpurun_out/r2a_bench_$w.err
done
timeout 120 python tools/pcie_bw.py > gpurun_out/r2a_pcie.log 2>&1; tail -5 gpurun_out/r2a_pcie.log
