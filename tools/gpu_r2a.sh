#!/bin/bash
# Round 2, first GPU pass: all GPU tests (no -x: collect every failure), then short benches over the engine options.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py build > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s $PYTEST_ARGS > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 40 gpurun_out/pytest_gpu.log | cut -c1-400
B="--steps 2 --warmup 1 --no-train --no-pano768 --no-cpu-baseline"
run() { echo "== bench.py $*"; timeout 600 python bench.py $B "$@" > gpurun_out/tmp.log 2> gpurun_out/tmp.err; echo "exit $?"; tail -n 1 gpurun_out/tmp.log | python tools/brief.py; tail -n 3 gpurun_out/tmp.err | cut -c1-300; cat gpurun_out/tmp.log >> gpurun_out/bench_all.log; }
run --streams 2 --fast-tail --profile-calls
run --streams 2 --ts-precision 1,1,1,1,1,1,1,3
