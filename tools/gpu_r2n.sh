#!/bin/bash
mkdir -p gpurun_out/ev
E=gpurun_out/ev
python __graft_entry__.py build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 5 gpurun_out/pytest_gpu.log | cut -c1-300
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 420 python tools/microbench.py --budget-s 60 > $E/microbench_fixed.jsonl 2> $E/microbench_fixed.err; echo "microbench fixed exit $?"
timeout 600 python tools/microbench.py --budget-s ${SWEEP_S:-200} --sweep > $E/microbench_sweep.jsonl 2> $E/microbench_sweep.err; echo "microbench sweep exit $?"
python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline --profile-calls 2>/dev/null | tail -1 > gpurun_out/bench_train.log; python tools/brief.py < gpurun_out/bench_train.log | head -4
