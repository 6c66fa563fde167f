#!/bin/bash
# ncu --set full of every kernel of ONE position group (tools/probes/one_group.py), exported to CSV on the box.
mkdir -p gpurun_out/prof
P=gpurun_out/prof
python __graft_entry__.py build > $P/build.log 2>&1
timeout 300 python tools/probes/one_group.py > $P/plain.log 2>&1 || { echo "plain run failed"; tail -5 $P/plain.log; exit 1; }
timeout 1200 ncu --set full --import-source on --clock-control none --nvtx --nvtx-include "prof/" -o /tmp/prof_group python tools/probes/one_group.py > $P/ncu_group.log 2>&1
echo "ncu exit $?"; tail -3 $P/ncu_group.log
ncu -i /tmp/prof_group.ncu-rep --page raw --csv > $P/group_raw.csv 2>/dev/null
python tools/ncu_summary.py $P/group_raw.csv > $P/group_summary.txt 2>&1
ls -la /tmp/prof_group.ncu-rep $P; wc -l $P/group_summary.txt
if [ "$KEEP_REP" == "1" ]; then cp /tmp/prof_group.ncu-rep $P/; fi
