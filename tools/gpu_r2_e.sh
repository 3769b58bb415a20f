#!/bin/bash
# round 2, GPU call E (1 GPU): tests, default bench, class capture with the bench line of the same run, sanitizers, memory ceilings
TAG=${1:-r2e}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=3 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
timeout 60 python tools/mem_bw.py > gpurun_out/${TAG}_mem_bw.json 2>&1; cat gpurun_out/${TAG}_mem_bw.json
timeout 500 ncu --set full --clock-control none -k 'regex:k_' -c 400 -o /tmp/prof_step \
    python bench.py --steps 1 --warmup 0 --projects 64 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu_step.log 2>&1
ncu -i /tmp/prof_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw_step.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_' -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu_launch.log 2>&1
( time timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_run.py ) > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_memcheck.log
( time SAN_SCALE=0.01 timeout 700 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_run.py ) > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_racecheck.log
du -sh gpurun_out
