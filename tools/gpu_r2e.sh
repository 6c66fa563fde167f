#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x -k "cta_pair" > gpurun_out/pytest_new.log 2>&1; echo "pair tests exit $?"; tail -n 8 gpurun_out/pytest_new.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-300
B="--steps 2 --warmup 1 --no-train --no-pano768 --no-cpu-baseline --no-strict"
run() { echo "== $PAIR bench.py $*"; timeout 600 python bench.py $B "$@" > gpurun_out/tmp.log 2> gpurun_out/tmp.err; echo "exit $?"; tail -n 1 gpurun_out/tmp.log | python tools/brief.py 2>/dev/null | head -${LINES_BRIEF:-30}; tail -n 3 gpurun_out/tmp.err | cut -c1-300; cat gpurun_out/tmp.log >> gpurun_out/bench_all.log; }
run --pair-mode 0 
run --pair-mode 1 
run --pair-mode 2
bash tools/gpu_ncu.sh
