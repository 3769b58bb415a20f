#!/bin/bash
# round 2, GPU call Q (1 GPU): inlined sums (A/B), chunks per block chosen per launch (configs[2] back to one chunk per
# block; e2e groups), full GPU test suite with the new parity tests
TAG=${1:-r2q}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x --durations=3 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
TDG_INLINE_SUMS=0 timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_noinline.json 2> gpurun_out/${TAG}_bench_c5_noinline.err
TDG_LAUNCH_WAVES=0 timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_waves0.json 2> gpurun_out/${TAG}_bench_c5_waves0.err
TDG_LAUNCH_WAVES=4 timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_waves4.json 2> gpurun_out/${TAG}_bench_c5_waves4.err
TDG_LAUNCH_WAVES=64 timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_waves64.json 2> gpurun_out/${TAG}_bench_c5_waves64.err
timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
du -sh gpurun_out
