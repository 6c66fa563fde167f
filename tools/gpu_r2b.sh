#!/bin/bash
# Round 2, re-entry pass: all GPU tests, then the default bench line and the reference arm, then a launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py build > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q $PYTEST_ARGS > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench default"; timeout 900 python bench.py --profile-calls > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "exit $?"
tail -n 1 gpurun_out/bench_default.log | python tools/brief.py; tail -n 3 gpurun_out/bench_default.err | cut -c1-300
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "exit $?"; tail -n 1 gpurun_out/bench_ref.log | cut -c1-600
