#!/bin/bash
# Full-size bench + ncu launch list + one full ncu capture of the tcgen05 GEMM (run via gpurun, 1 GPU).
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_full.log
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 21780 -c 7260 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 90 -c 6 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
