#!/bin/bash
# What the driver runs at round end, in the same order: GPU tests, smoke, [the reference arm,] the default bench.
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/smoke.log
if [ "$1" == "reference" ]; then
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "reference exit $?"; tail -n 1 gpurun_out/bench_reference.log | cut -c1-400
fi
timeout 300 python bench.py > gpurun_out/bench_full.log 2>gpurun_out/bench_full.err; echo "bench exit $?"; tail -n 1 gpurun_out/bench_full.log | cut -c1-400; tail -n 3 gpurun_out/bench_full.err
