#!/bin/bash
# ncu evidence for profiles/ (1 GPU).  Launch lists (gpu__time_duration) of one timed step of each bench workload, and
# `--set full` captures of the dominant kernels, exported to CSV on the box (the .ncu-rep files are too big to travel).
# Each ncu command runs only after the same command exited 0 without ncu.  `--warmup 1 --skip-e2e` only shortens the
# serialised run under ncu; the workload, batch and kernels are those of the default bench command.
mkdir -p gpurun_out/prof
P=gpurun_out/prof
python __graft_entry__.py build > $P/build.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline"
timeout 600 $CMD > $P/plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 7400 -c 7400 --csv --log-file $P/launches_pano.csv $CMD > $P/ncu_launches.log 2>&1
echo "ncu launches exit $?"; wc -l $P/launches_pano.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel|upblur_act|sphere_pack_shared|pack_act_kernel|conv_small_cout" -s 870 -c 24 -o /tmp/prof_pano $CMD > $P/ncu_full_pano.log 2>&1
echo "ncu full pano exit $?"
ncu -i /tmp/prof_pano.ncu-rep --page details --csv > $P/full_pano_details.csv 2>/dev/null
ncu -i /tmp/prof_pano.ncu-rep --page raw --csv > $P/full_pano_raw.csv 2>/dev/null
CMDT="python bench.py --workload train --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline"
timeout 600 $CMDT > $P/plain_train.log 2>&1 || { echo "plain train failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6600 -c 7000 --csv --log-file $P/launches_train.csv $CMDT > $P/ncu_launches_train.log 2>&1
echo "ncu train launches exit $?"; wc -l $P/launches_train.csv
timeout 600 ncu --set full --clock-control none -k regex:"conv_wgrad_gemm_kernel" -s 40 -c 6 -o /tmp/prof_wgrad $CMDT > $P/ncu_full_wgrad.log 2>&1
echo "ncu full wgrad exit $?"
ncu -i /tmp/prof_wgrad.ncu-rep --page details --csv > $P/full_wgrad_details.csv 2>/dev/null
ncu -i /tmp/prof_wgrad.ncu-rep --page raw --csv > $P/full_wgrad_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la $P
