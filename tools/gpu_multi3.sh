#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --workload train --no-cpu-baseline > gpurun_out/train_n$N.log 2>gpurun_out/train_n$N.err; echo "train N=$N exit $?"
tail -n 1 gpurun_out/train_n$N.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value %.1f img/s ms/step %.2f parts %s in_sync %s graphs %s' % (d['value'], d['ms_per_step'], {k: round(v, 2) for k, v in d['config']['part_ms'].items()}, d['config']['replicas_in_sync_after_run'], d['config']['cuda_graphs']))"
grep -i "graph capture failed\|Error\|error" gpurun_out/train_n$N.err | head -5
timeout 300 python bench.py --workload train --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N=1 value %.1f img/s ms/step %.2f' % (d['value'], d['ms_per_step']))"
