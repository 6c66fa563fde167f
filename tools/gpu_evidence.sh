#!/bin/bash
# Round-2 evidence for profiles/: microbenchmarks with clock records, ncu launch lists of one panorama step and one training
# iteration (eager launches on one stream, inside the NVTX range bench.py opens around its timed region), and `ncu --set full`
# of every tcgen05 GEMM launch of one position group.  Each ncu command runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out/ev
E=gpurun_out/ev
python __graft_entry__.py build > $E/build.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $E/gpu.txt 2>&1
timeout 420 python tools/microbench.py --budget-s 60 > $E/microbench_fixed.jsonl 2> $E/microbench_fixed.err; echo "microbench fixed exit $?"
timeout 600 python tools/microbench.py --budget-s ${SWEEP_S:-240} --sweep > $E/microbench_sweep.jsonl 2> $E/microbench_sweep.err; echo "microbench sweep exit $?"
PANO="python bench.py --steps 1 --warmup 1 --skip-e2e --skip-profile --no-cpu-baseline --no-strict --no-train --no-pano768 --no-graphs --streams 1"
timeout 600 $PANO > $E/pano_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spgan_timed" --csv --log-file $E/launches_pano.csv $PANO > $E/ncu_launches_pano.log 2>&1
echo "pano launch list exit $?"; wc -l $E/launches_pano.csv
TRAIN="python bench.py --workload train --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline --no-graphs"
timeout 600 $TRAIN > $E/train_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "spgan_timed" --csv --log-file $E/launches_train.csv $TRAIN > $E/ncu_launches_train.log 2>&1
echo "train launch list exit $?"; wc -l $E/launches_train.csv
timeout 300 python tools/probes/one_group.py > $E/group_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --nvtx --nvtx-include "prof/" -k regex:"conv_gemm|sphere_pack_v3|upblur_pack|coord_taps|concat_repack" -o /tmp/ev_group python tools/probes/one_group.py > $E/ncu_group.log 2>&1
echo "ncu full exit $?"
ncu -i /tmp/ev_group.ncu-rep --page raw --csv > $E/group_raw.csv 2>/dev/null
python tools/ncu_summary.py $E/group_raw.csv > $E/group_summary.txt 2>&1
rm -f $E/group_raw.csv.tmp; ls -la $E; du -sh gpurun_out
