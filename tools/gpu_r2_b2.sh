#!/bin/bash
# round 2, GPU call B2 (1 GPU): device-built segment tables + r1 voice kernel body: tests, default bench (with secondary), profiles
TAG=${1:-r2b2}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=3 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
TDG_FAST_SIN=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_c5_fastsin.json 2> gpurun_out/${TAG}_bench_c5_fastsin.err
PASS='regex:k_(fx|voice|bp_|norm_|sample_|quant|resample|chunk_ranges|env_table)'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$PASS" -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k_voice' -c 6 -o /tmp/prof_voice \
    python bench.py --steps 1 --warmup 1 --projects 64 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu_voice.log 2>&1
ncu -i /tmp/prof_voice.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw_voice.csv 2>/dev/null
ncu -i /tmp/prof_voice.ncu-rep --page source --print-source sass --csv > gpurun_out/${TAG}_prof_source_voice.csv 2>gpurun_out/${TAG}_prof_source_voice.err
timeout 400 ncu --set full --clock-control none -k 'regex:k_fx|k_bp_apply|k_bp_aggregate|k_sample_|k_quantise|k_chunk_ranges|k_voice_seg' -c 40 -o /tmp/prof_rest \
    python bench.py --steps 1 --warmup 0 --projects 64 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu_rest.log 2>&1
ncu -i /tmp/prof_rest.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw_rest.csv 2>/dev/null
ls -la /tmp/*.ncu-rep; du -sh gpurun_out
