#!/bin/bash
# usage: run_gpu.sh <timeout> '<command>' — builds first and refuses to spend GPU time on a broken build
set -e
cd /root/repo
python termdaw_b200/csrc/build.py > /tmp/build.log 2>&1 || { tail -20 /tmp/build.log; echo "BUILD FAILED"; exit 1; }
make -C oracle > /dev/null
python -c "import termdaw_b200; termdaw_b200.load_library()" || { echo "LOAD FAILED"; exit 1; }
python -m pytest tests -q -x -m "not gpu" 2>&1 | tail -2
/usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
