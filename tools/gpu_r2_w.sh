#!/bin/bash
# round 2, GPU call W (1 GPU): refresh pipelined inside the batched call (tdg_refresh_normalize_render_many): parity + e2e A/B
TAG=${1:-r2w}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_planner_invariants.py tests/test_gpu_parity.py tests/test_pass_reuse.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest_gpu_subset.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu_subset.log
B="--steps 8 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
TDG_E2E_FUSED_REFRESH=0 timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_rewind_first.json 2> gpurun_out/${TAG}_bench_c5_rewind_first.err
