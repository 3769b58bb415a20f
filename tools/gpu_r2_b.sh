#!/bin/bash
# round 2, GPU call B (1 GPU): device-side tables + exact sin — full -m gpu suite, bench c5 / c3 / c4, ncu launch list,
# --set full of the heaviest kernels with source pages for the voice kernels
TAG=${1:-r2b}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
( time timeout 400 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
TDG_FAST_SIN=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c5_fastsin.json 2> gpurun_out/${TAG}_bench_c5_fastsin.err
for w in c3 c4; do
  ( time timeout 420 python bench.py --workload $w --steps 3 --warmup 2 ) > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "$w rc=$?"; tail -c 300 gpurun_out/${TAG}_bench_$w.err
done
PASS='regex:k_(fx|voice|bp_|norm_|sample_|quant|resample|chunk_ranges|env_table)'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$PASS" -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_voice|k_fx|k_bp_apply|k_bp_aggregate|k_sample_|k_quantise|k_chunk_ranges' -s 40 -c 40 -o /tmp/prof_top \
    python bench.py --steps 1 --warmup 1 --projects 64 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/prof_top.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
ncu -i /tmp/prof_top.ncu-rep --page source --csv -k regex:k_voice > gpurun_out/${TAG}_prof_source_voice.csv 2>/dev/null
ls -la /tmp/prof_top.ncu-rep; du -sh gpurun_out
