#!/bin/bash
# End-of-round run (second session set of round 2): all GPU tests, smoke, the default bench line, microbenchmarks (fixed cases
# with the legacy A/B lines + the configs[4] sweep), ncu launch lists of one panorama step and one training iteration, and
# `ncu --set full` of the streamed HBM kernels.  Each ncu command runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out/ev
E=gpurun_out/ev
python __graft_entry__.py build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $E/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-300
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --profile-calls > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -n 1 gpurun_out/bench_default.log | python tools/brief.py 2>/dev/null | head -8
timeout 420 python tools/microbench.py --budget-s 100 --legacy-ab > $E/microbench_fixed.jsonl 2> $E/microbench_fixed.err; echo "microbench fixed exit $?"
if [ "${SKIP_SWEEP:-0}" != "1" ]; then
timeout 600 python tools/microbench.py --budget-s ${SWEEP_S:-240} --sweep > $E/microbench_sweep.jsonl 2> $E/microbench_sweep.err; echo "microbench sweep exit $?"
fi
timeout 120 python tools/probes/hbm_kernels.py > $E/hbm_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none -k regex:"fir_stream|gather_stream" -c 10 -o /tmp/hbm python tools/probes/hbm_kernels.py > $E/ncu_hbm.log 2>&1
echo "ncu hbm exit $?"
ncu -i /tmp/hbm.ncu-rep --page raw --csv > $E/hbm_raw.csv 2>/dev/null
python tools/ncu_summary.py $E/hbm_raw.csv > $E/hbm_summary.txt 2>&1; rm -f $E/hbm_raw.csv
if [ "${SKIP_LISTS:-0}" != "1" ]; then
PANO="python bench.py --steps 1 --warmup 1 --skip-e2e --skip-profile --no-cpu-baseline --no-strict --no-train --no-pano768 --no-graphs --streams 1"
timeout 600 $PANO > $E/pano_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spgan_timed" --csv --log-file $E/launches_pano.csv $PANO > $E/ncu_launches_pano.log 2>&1
echo "pano launch list exit $?"; wc -l $E/launches_pano.csv
TRAIN="python bench.py --workload train --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline --no-graphs"
timeout 600 $TRAIN > $E/train_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spgan_timed" --csv --log-file $E/launches_train.csv $TRAIN > $E/ncu_launches_train.log 2>&1
echo "train launch list exit $?"; wc -l $E/launches_train.csv; tail -2 $E/ncu_launches_train.log
fi
