#!/bin/bash
# round 2, GPU call R (1 GPU): the round's evidence run on the tuned kernels — full GPU test suite, the default bench line,
# the reference arm, configs[0..3] lines, per-class ncu metrics of one step + the bench line of that same run, the ncu launch
# list of the bench command, --set full capture of the heaviest kernels, sanitizers, the e2e job timeline
TAG=${1:-r2r}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=3 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err
for w in c1 c2; do timeout 200 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; done
PASS='regex:k_(fx|voice|bp_|norm_|sample_|quant|resample|chunk_ranges|seg_|env_table)'
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
timeout 600 ncu --metrics $M --clock-control none -k "$PASS" -c 400 --csv --log-file gpurun_out/${TAG}_step_metrics.csv \
    python bench.py --steps 1 --warmup 0 --projects 64 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_step_metrics_bench.json 2> gpurun_out/${TAG}_step_metrics_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$PASS" -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu_launch.log 2>&1
TDG_INLINE_SUMS=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_c5_noinline.json 2> gpurun_out/${TAG}_bench_c5_noinline.err
( time SAN_SCALE=0.1 timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_run.py ) > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_memcheck.log
( time SAN_SCALE=0.02 timeout 700 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_run.py ) > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_racecheck.log
TDG_TRACE=1 timeout 200 python tools/trace_e2e.py 512 > /dev/null 2> gpurun_out/${TAG}_e2e_timeline.log
du -sh gpurun_out
