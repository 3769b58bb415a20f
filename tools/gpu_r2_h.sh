#!/bin/bash
# round 2, GPU call H (1 GPU): resampler staging + single-chain sinf on the device (parity, bench), fused band pass A/B,
# --set full capture (with source) of the fused band pass / adsr / voice kernels, e2e timeline at N=1, configs[0] bench
TAG=${1:-r2h}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=3 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
TDG_BP_FUSED=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_c5_bpfused.json 2> gpurun_out/${TAG}_bench_c5_bpfused.err
timeout 200 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c1.json 2> gpurun_out/${TAG}_bench_c1.err
TDG_TRACE=1 timeout 200 python tools/trace_e2e.py 512 > /dev/null 2> gpurun_out/${TAG}_e2e_timeline.log
TDG_BP_FUSED=1 timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_bp_fused|k_fx|k_voice' -c 24 -o /tmp/prof_h \
    python bench.py --steps 1 --warmup 0 --projects 64 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/prof_h.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
python tools/ncu_stalls.py gpurun_out/${TAG}_prof_raw.csv > gpurun_out/${TAG}_stalls.txt 2>&1
ncu -i /tmp/prof_h.ncu-rep --page source --csv -k regex:k_bp_fused > gpurun_out/${TAG}_src_bp_fused.csv 2>/dev/null
ls -la /tmp/prof_h.ncu-rep; du -sh gpurun_out
