#!/bin/bash
# round 2, GPU call L (1 GPU): defaults after call K (voices 4 chunks per block, adsr 8, quantise 4, fused band pass in 128-thread
# blocks, no length limit); variants: voices 8 / 16, band pass 64 threads, adsr 16, gathers 16
TAG=${1:-r2l}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_pass_reuse.py tests/test_playback.py tests/test_golden.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest_gpu_subset.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu_subset.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
cp termdaw_b200/csrc/libtermdaw_b200.so /tmp/main.so
for v in vc8 vc16 bp64 ac16 g16; do
  cp termdaw_b200/csrc/variants/lib_$v.so termdaw_b200/csrc/libtermdaw_b200.so
  timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_$v.json 2> gpurun_out/${TAG}_bench_c5_$v.err
done
cp /tmp/main.so termdaw_b200/csrc/libtermdaw_b200.so
timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
du -sh gpurun_out
