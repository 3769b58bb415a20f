#!/bin/bash
# run ON the GPU box (under gpurun): bench + ncu launch list + ncu full capture of every pass kernel
# usage: tools_profile.sh <tag>
TAG=${1:-s1}
PASS='regex:k_(fx|voice|bp_|norm_|sample_|quant|resample|scan)'
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 400 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -k "$PASS" -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "$PASS" -s 80 -c 80 -o /tmp/prof_all \
    python bench.py --steps 1 --warmup 1 --projects 64 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_raw_$TAG.csv 2>/dev/null
ls -la /tmp/prof_all.ncu-rep
S=$(stat -c %s /tmp/prof_all.ncu-rep 2>/dev/null || echo 0)
if [ "$S" -lt 45000000 ]; then cp /tmp/prof_all.ncu-rep gpurun_out/prof_all_$TAG.ncu-rep; fi
du -sh gpurun_out
