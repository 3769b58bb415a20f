#!/bin/bash
# round 2, GPU call X (1 GPU): full GPU test suite and the default bench line at the final commit
TAG=${1:-r2x}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=3 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
