#!/bin/bash
# round 2, GPU call J (1 GPU): fused band pass with cp.async staging (now the default), gather kernels walking 4 chunks per
# block (A/B against 1 and 8), k_fx<adsr> chunks-per-block variants; --set full capture of the changed kernels
TAG=${1:-r2j}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x --durations=3 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
TDG_BP_FUSED=0 timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_bp3.json 2> gpurun_out/${TAG}_bench_c5_bp3.err
cp termdaw_b200/csrc/libtermdaw_b200.so /tmp/main.so
for v in g1 g8 ac2 ac4; do
  cp termdaw_b200/csrc/variants/lib_$v.so termdaw_b200/csrc/libtermdaw_b200.so
  timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5_$v.json 2> gpurun_out/${TAG}_bench_c5_$v.err
done
cp /tmp/main.so termdaw_b200/csrc/libtermdaw_b200.so
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:k_bp_fused|k_fx|k_sample' -c 40 -o /tmp/prof_j \
    python bench.py --steps 1 --warmup 0 --projects 64 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/prof_j.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
python tools/ncu_stalls.py gpurun_out/${TAG}_prof_raw.csv > gpurun_out/${TAG}_stalls.txt 2>&1
ls -la /tmp/prof_j.ncu-rep; du -sh gpurun_out
