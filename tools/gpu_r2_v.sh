#!/bin/bash
# round 2, GPU call V (1 GPU): chunks per block of the grouped k_fx<sum|norm_peak> instantiations: 4 (default) vs 1 / 2 / 8
TAG=${1:-r2v}
mkdir -p gpurun_out
B="--steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
cp termdaw_b200/csrc/libtermdaw_b200.so /tmp/main.so
for v in fxgc1 fxgc2 fxgc8; do
  cp termdaw_b200/csrc/variants/lib_$v.so termdaw_b200/csrc/libtermdaw_b200.so
  timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_$v.json 2> gpurun_out/${TAG}_bench_c5_$v.err
done
cp /tmp/main.so termdaw_b200/csrc/libtermdaw_b200.so
