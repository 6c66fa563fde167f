#!/bin/bash
# One iteration on the GPU box: full GPU test suite, smoke, full-size bench, launch list of one bench run.
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_full.log
if [ "$1" == "ncu" ]; then
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 9000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
fi
