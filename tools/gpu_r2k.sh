#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-300
python bench.py --steps 3 --warmup 3 --profile-calls > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench exit $?"
tail -n 1 gpurun_out/bench_default.log | python tools/brief.py 2>/dev/null | head -40
