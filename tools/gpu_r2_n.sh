#!/bin/bash
# round 2, GPU call N (1 GPU): full GPU test suite on the tuned kernels (wavetable last-row shortcut: 512-voice parity test),
# occupancy variants of the voice kernels, chunks-per-block for k_fx<sum|norm_peak>, configs[2], pipeline group 32
TAG=${1:-r2n}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x --durations=3 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
cp termdaw_b200/csrc/libtermdaw_b200.so /tmp/main.so
for v in sm4 sm6 wm4 wm6 fx2 fx4; do
  cp termdaw_b200/csrc/variants/lib_$v.so termdaw_b200/csrc/libtermdaw_b200.so
  timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_$v.json 2> gpurun_out/${TAG}_bench_c5_$v.err
done
cp /tmp/main.so termdaw_b200/csrc/libtermdaw_b200.so
du -sh gpurun_out
