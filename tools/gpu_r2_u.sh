#!/bin/bash
# round 2, GPU call U (1 GPU): grouped k_fx with the expanded pointer list staged in shared memory (A/B: inlining off, 4 blocks per SM)
TAG=${1:-r2u}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_planner_invariants.py tests/test_gpu_parity.py tests/test_pass_reuse.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest_gpu_subset.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu_subset.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
TDG_INLINE_SUMS=0 timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_noinline.json 2> gpurun_out/${TAG}_bench_c5_noinline.err
cp termdaw_b200/csrc/libtermdaw_b200.so /tmp/main.so
for v in fxg4; do
  cp termdaw_b200/csrc/variants/lib_$v.so termdaw_b200/csrc/libtermdaw_b200.so
  timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_$v.json 2> gpurun_out/${TAG}_bench_c5_$v.err
done
cp /tmp/main.so termdaw_b200/csrc/libtermdaw_b200.so
