#!/bin/bash
mkdir -p gpurun_out/ev
E=gpurun_out/ev
python __graft_entry__.py build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 420 python tools/microbench.py --budget-s 80 > $E/microbench_fixed.jsonl 2> $E/microbench_fixed.err; echo "microbench fixed exit $?"; tail -3 $E/microbench_fixed.err
timeout 600 python tools/microbench.py --budget-s ${SWEEP_S:-220} --sweep > $E/microbench_sweep.jsonl 2> $E/microbench_sweep.err; echo "microbench sweep exit $?"; tail -3 $E/microbench_sweep.err
wc -l $E/*.jsonl
