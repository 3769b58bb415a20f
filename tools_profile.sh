#!/bin/bash
# run ON the GPU box (under gpurun): bench + ncu launch list + ncu full capture of every pass kernel
# usage: tools_profile.sh <tag>
TAG=${1:-s1}
PASS='regex:k_(fx|voice|bp_|norm_|sample_|quant|resample|scan)'
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 400 gpurun_out/bench_$TAG.err
[ -n "$SKIP_REF" ] || python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1
# launch list of the SAME command line as the bench (per-launch device time, cold cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$PASS" -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
# full capture: one timed step at 64 projects (same kernels, smaller grids so that 40 replays per launch stay short)
timeout 900 ncu --set full --clock-control none -k "$PASS" -s 80 -c 80 -o /tmp/prof_all \
    python bench.py --steps 1 --warmup 1 --projects 64 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_raw_$TAG.csv 2>/dev/null
ls -la /tmp/prof_all.ncu-rep
du -sh gpurun_out
