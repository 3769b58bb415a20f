#!/bin/bash
# run ON the GPU box (under gpurun): the short profile pass — bench (+ an A/B line with pass reuse off), the ncu
# launch list of the same command, and a --set full capture of the few heaviest kernels only
# usage: tools_profile_lite.sh <tag>
TAG=${1:-s6}
PASS='regex:k_(fx|voice|bp_|norm_|sample_|quant|resample|scan)'
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 300 gpurun_out/bench_$TAG.err
TDG_PASS_REUSE=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_noreuse.json 2> gpurun_out/bench_${TAG}_noreuse.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k "$PASS" -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
# full capture: the voice kernels and the adsr fx kernel of one 64-project step (8 launches in all)
timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:k_voice|k_bp_apply' -s 8 -c 8 -o /tmp/prof_top \
    python bench.py --steps 1 --warmup 1 --projects 64 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i /tmp/prof_top.ncu-rep --page raw --csv > gpurun_out/prof_raw_$TAG.csv 2>/dev/null
ls -la /tmp/prof_top.ncu-rep; du -sh gpurun_out
