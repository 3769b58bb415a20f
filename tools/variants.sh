python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in m5 m6; do TDG_TEST_LIB=$PWD/termdaw_b200/csrc/lib_$v.so python bench.py --steps 2 --warmup 3 --no-cpu-baseline --projects 256 > gpurun_out/bench_v_$v.json 2> gpurun_out/bench_v_$v.err; done; python bench.py --steps 2 --warmup 3 --no-cpu-baseline --projects 256 > gpurun_out/bench_v_base2.json 2> gpurun_out/bench_v_base2.err
