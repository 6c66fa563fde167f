#!/bin/bash
# ncu evidence for profiles/ (1 GPU).  Launch list (gpu__time_duration) of one timed step of the default bench workload and
# `--set full` captures of the dominant kernels, exported to CSV on the box (the .ncu-rep files are too big to travel).
# Each ncu command runs only after the same command exited 0 without ncu.  `--warmup 1 --skip-e2e` only shortens the
# serialised run under ncu; the workload, batch and kernels are those of the default bench command.
mkdir -p gpurun_out/prof
P=gpurun_out/prof
python __graft_entry__.py build > $P/build.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline"
timeout 600 $CMD > $P/plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
timeout 600 ncu --set full --clock-control none -k regex:"conv_gemm_kernel|upblur_act|sphere_pack_shared|pack_act_kernel" -s 735 -c 20 -o /tmp/prof_pano $CMD > $P/ncu_full_pano.log 2>&1
echo "ncu full pano exit $?"
ncu -i /tmp/prof_pano.ncu-rep --page raw --csv > $P/full_pano_raw.csv 2>/dev/null
if [ "$1" == "launches" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 6000 --csv --log-file $P/launches_pano.csv $CMD > $P/ncu_launches.log 2>&1
echo "ncu launches exit $?"; wc -l $P/launches_pano.csv
fi
du -sh gpurun_out; ls -la $P
