#!/bin/bash
# round 2, GPU call G (1 GPU): fused band pass with interleaved filters (A/B), e2e group sweep, whole-step metrics capture + its bench line
TAG=${1:-r2g}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=3 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
( time TDG_BP_FUSED=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_pass_reuse.py tests/test_playback.py -m gpu -q ) > gpurun_out/${TAG}_pytest_gpu_bpfused.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu_bpfused.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
TDG_BP_FUSED=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_c5_bpfused.json 2> gpurun_out/${TAG}_bench_c5_bpfused.err
for g in 32 128 256; do TDG_E2E_GROUP=$g timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_c5_group$g.json 2> gpurun_out/${TAG}_bench_c5_group$g.err; done
PASS='regex:k_(fx|voice|bp_|norm_|sample_|quant|resample|chunk_ranges|seg_|env_table)'
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
timeout 600 ncu --metrics $M --clock-control none -k "$PASS" -c 400 --csv --log-file gpurun_out/${TAG}_step_metrics.csv \
    python bench.py --steps 1 --warmup 0 --projects 64 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_step_metrics_bench.json 2> gpurun_out/${TAG}_step_metrics_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$PASS" -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu_launch.log 2>&1
du -sh gpurun_out
