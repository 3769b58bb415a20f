#!/bin/bash
# round 2, GPU call I (1 GPU): strip-based fused band pass (parity with TDG_BP_FUSED=1, A/B bench), k_fx<adsr> with two runs
# per thread (A/B against block-size / occupancy variants), --set full capture of the streaming kernels
TAG=${1:-r2i}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x --durations=3 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
( time TDG_BP_FUSED=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_pass_reuse.py tests/test_playback.py tests/test_golden.py -m gpu -q ) > gpurun_out/${TAG}_pytest_gpu_bpfused.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu_bpfused.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
TDG_BP_FUSED=1 timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_bpfused.json 2> gpurun_out/${TAG}_bench_c5_bpfused.err
TDG_BP_FUSED=1 timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2_bpfused.json 2> gpurun_out/${TAG}_bench_c2_bpfused.err
cp termdaw_b200/csrc/libtermdaw_b200.so /tmp/main.so
for v in a7 a256 a256m5; do
  cp termdaw_b200/csrc/variants/lib_$v.so termdaw_b200/csrc/libtermdaw_b200.so
  timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_$v.json 2> gpurun_out/${TAG}_bench_c5_$v.err
done
cp /tmp/main.so termdaw_b200/csrc/libtermdaw_b200.so
TDG_BP_FUSED=1 timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:k_bp_fused|k_fx|k_sample|k_quantise' -c 40 -o /tmp/prof_i \
    python bench.py --steps 1 --warmup 0 --projects 64 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/prof_i.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
python tools/ncu_stalls.py gpurun_out/${TAG}_prof_raw.csv > gpurun_out/${TAG}_stalls.txt 2>&1
ls -la /tmp/prof_i.ncu-rep; du -sh gpurun_out
