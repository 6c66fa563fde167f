#!/bin/bash
# End-of-round run: all GPU tests, smoke, the default bench line, microbenchmarks (fixed cases + sweep) and the train launch list.
mkdir -p gpurun_out/ev
E=gpurun_out/ev
python __graft_entry__.py build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-300
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --profile-calls > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -n 1 gpurun_out/bench_default.log | python tools/brief.py 2>/dev/null | head -8
timeout 420 python tools/microbench.py --budget-s 80 > $E/microbench_fixed.jsonl 2> $E/microbench_fixed.err; echo "microbench fixed exit $?"
timeout 600 python tools/microbench.py --budget-s ${SWEEP_S:-220} --sweep > $E/microbench_sweep.jsonl 2> $E/microbench_sweep.err; echo "microbench sweep exit $?"
TRAIN="python bench.py --workload train --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline --no-graphs"
timeout 600 $TRAIN > $E/train_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spgan_timed" --csv --log-file $E/launches_train.csv $TRAIN > $E/ncu_launches_train.log 2>&1
echo "train launch list exit $?"; wc -l $E/launches_train.csv; tail -2 $E/ncu_launches_train.log
