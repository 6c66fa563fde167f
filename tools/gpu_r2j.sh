#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests -m gpu -q -x -s -k "structure_chain or ema_multi" 2>&1 | grep -E "structure chain|passed|failed|Error|error|assert" | cut -c1-300 | tail -20
python tools/probes/pair_ab.py 2>&1 | tail -32
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-300
B="--steps 2 --warmup 1 --no-train --no-pano768 --no-cpu-baseline --no-strict"
run() { echo "== bench.py $*"; timeout 600 python bench.py $B "$@" > gpurun_out/tmp.log 2> gpurun_out/tmp.err; echo "exit $?"; tail -n 1 gpurun_out/tmp.log | python tools/brief.py 2>/dev/null | head -${LINES_BRIEF:-30}; tail -n 3 gpurun_out/tmp.err | cut -c1-300; cat gpurun_out/tmp.log >> gpurun_out/bench_all.log; }
run --profile-calls
