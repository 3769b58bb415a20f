#!/bin/bash
# round 2, GPU call K (1 GPU): chunks-per-block defaults (gathers 8, adsr 4), variants: adsr 8, fused band pass in
# 128-thread blocks, voice kernels 2 / 4 chunks per block, quantise 4; configs[3] with the fused band pass on an hour-long pass
TAG=${1:-r2k}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x --durations=3 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
TDG_BP_FUSED_MAX_TILES=262140 timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4_fusedlong.json 2> gpurun_out/${TAG}_bench_c4_fusedlong.err
cp termdaw_b200/csrc/libtermdaw_b200.so /tmp/main.so
for v in ac8 bp128 vc2 vc4 q4; do
  cp termdaw_b200/csrc/variants/lib_$v.so termdaw_b200/csrc/libtermdaw_b200.so
  timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_$v.json 2> gpurun_out/${TAG}_bench_c5_$v.err
done
cp /tmp/main.so termdaw_b200/csrc/libtermdaw_b200.so
du -sh gpurun_out
