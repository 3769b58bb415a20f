#!/bin/bash
# round 2, GPU call F (1 GPU): fused single-pass band pass (A/B), class capture with the bench line of the same run, c1 / c2 lines
TAG=${1:-r2f}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=3 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
TDG_BP_FUSED=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_c5_bp3.json 2> gpurun_out/${TAG}_bench_c5_bp3.err
TDG_FAST_SIN=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_c5_fastsin.json 2> gpurun_out/${TAG}_bench_c5_fastsin.err
for w in c1 c2; do timeout 200 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; done
PASS='regex:k_(fx|voice|bp_|norm_|sample_|quant|resample|chunk_ranges|seg_|env_table)'
timeout 500 ncu --set full --clock-control none -k "$PASS" -c 200 -o /tmp/prof_step \
    python bench.py --steps 1 --warmup 0 --projects 64 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu_step.log 2>&1
ncu -i /tmp/prof_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw_step.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$PASS" -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 200 ncu --set full --clock-control none -k 'regex:k_resample' -c 4 -o /tmp/prof_rs python bench.py --workload c1 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c1.log 2>&1
ncu -i /tmp/prof_rs.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw_c1.csv 2>/dev/null
du -sh gpurun_out
