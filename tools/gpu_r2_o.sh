#!/bin/bash
# round 2, GPU call O (1 GPU): end-to-end pipeline sweep — lanes (2, 3, 4) x group size x taper on/off
TAG=${1:-r2o}
mkdir -p gpurun_out
B="--steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.err
for lanes in 2 3 4; do for taper in 1 0; do
  TDG_LANES=$lanes TDG_TAPER=$taper TDG_E2E_GROUPS="32:0.8,16:0.8,24:0.8,48:0.8" timeout 200 python tools/trace_e2e.py 512 > /dev/null 2> gpurun_out/${TAG}_e2e_sweep_l${lanes}_t${taper}.log
  echo "== lanes $lanes taper $taper"; grep "combined it [23]" gpurun_out/${TAG}_e2e_sweep_l${lanes}_t${taper}.log
done; done
du -sh gpurun_out
