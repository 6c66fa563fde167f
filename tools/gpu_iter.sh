#!/bin/bash
# One iteration on the GPU box: full GPU test suite, smoke, full-size benches (panorama + train); with "ncu" also the
# launch lists of one bench run of each workload.
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_full.log
timeout 900 python bench.py --workload train > gpurun_out/bench_train.log 2>&1; echo "bench train exit $?"; tail -n 3 gpurun_out/bench_train.log
if [ "$1" == "ncu" ]; then
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 9000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
CMD="python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_train.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 25000 --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/ncu_launches_train.log 2>&1
echo "ncu train launches exit $?"
fi
