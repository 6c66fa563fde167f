#!/bin/bash
# Quick GPU iteration: GPU tests, then both benches with the per-entry-point event profile (no ncu).
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x $PYTEST_ARGS > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --profile-calls --no-cpu-baseline --no-train > gpurun_out/bench_prof.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_prof.log
timeout 600 python bench.py --workload train --profile-calls --no-cpu-baseline > gpurun_out/bench_train_prof.log 2>gpurun_out/bench_train_err.log; echo "bench train exit $?"; tail -n 1 gpurun_out/bench_train_prof.log | cut -c1-300; tail -n 3 gpurun_out/bench_train_err.log
for extra in "$@"; do
  echo "== bench.py $extra"; timeout 600 python bench.py $extra --no-cpu-baseline 2>&1 | tail -n 1
done
