#!/bin/bash
# round 2, GPU call M (1 GPU): voices 8 chunks per block + hoisted offsets (A/B), end-to-end pipeline sweep (group size,
# tail ratio, lanes) now that the kernels are faster than the PCM copy
TAG=${1:-r2m}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_pass_reuse.py tests/test_playback.py tests/test_golden.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest_gpu_subset.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu_subset.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
cp termdaw_b200/csrc/libtermdaw_b200.so /tmp/main.so
for v in nohoist; do
  cp termdaw_b200/csrc/variants/lib_$v.so termdaw_b200/csrc/libtermdaw_b200.so
  timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_$v.json 2> gpurun_out/${TAG}_bench_c5_$v.err
done
cp /tmp/main.so termdaw_b200/csrc/libtermdaw_b200.so
TDG_E2E_GROUPS="64:0.8,64:0.5,64:0.1,32:0.8,32:0.5,32:0.1,128:0.5,16:0.5" timeout 300 python tools/trace_e2e.py 512 > /dev/null 2> gpurun_out/${TAG}_e2e_sweep_lanes2.log
TDG_LANES=1 TDG_E2E_GROUPS="64:0.8,64:0.1,32:0.8,32:0.1,16:0.5" timeout 300 python tools/trace_e2e.py 512 > /dev/null 2> gpurun_out/${TAG}_e2e_sweep_lanes1.log
TDG_TRACE=1 TDG_E2E_GROUPS="64:0.8" timeout 200 python tools/trace_e2e.py 512 > /dev/null 2> gpurun_out/${TAG}_e2e_timeline.log
grep "combined it [23]" gpurun_out/${TAG}_e2e_sweep_lanes2.log gpurun_out/${TAG}_e2e_sweep_lanes1.log
du -sh gpurun_out
