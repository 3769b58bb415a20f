#!/bin/bash
# round 2, GPU call Y (1 GPU): programmatic dependent launch of the pass kernels (TDG_PDL=1): full parity suite + A/B bench
TAG=${1:-r2y}
mkdir -p gpurun_out
( time TDG_PDL=1 timeout 900 python -m pytest tests -m gpu -q -x --durations=3 ) > gpurun_out/${TAG}_pytest_gpu_pdl.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu_pdl.log
B="--steps 8 --warmup 3 --no-cpu-baseline --no-secondary"
TDG_PDL=1 timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_pdl.json 2> gpurun_out/${TAG}_bench_c5_pdl.err; tail -c 300 gpurun_out/${TAG}_bench_c5_pdl.err
TDG_PDL=0 timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_nopdl.json 2> gpurun_out/${TAG}_bench_c5_nopdl.err
TDG_PDL=1 timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_pdl2.json 2> gpurun_out/${TAG}_bench_c5_pdl2.err
for w in c1 c2; do TDG_PDL=1 timeout 200 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_${w}_pdl.json 2> gpurun_out/${TAG}_bench_${w}_pdl.err; done
( time TDG_PDL=1 SAN_SCALE=0.02 timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_run.py ) > gpurun_out/${TAG}_sanitizer_racecheck_pdl.log 2>&1; echo "racecheck rc=$?"
( time TDG_PDL=1 SAN_SCALE=0.05 timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_run.py ) > gpurun_out/${TAG}_sanitizer_memcheck_pdl.log 2>&1; echo "memcheck rc=$?"
