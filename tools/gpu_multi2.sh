#!/bin/bash
# N-GPU validation (N = $1): the default bench line under torchrun (weak-scaling panoramas + strict + 768x1536 lattice-sharded
# + the data-parallel train object) and the reference arm's non-zero ranks.
N=${1:-2}
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 1200 $TR bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_n$N.log 2>gpurun_out/bench_n$N.err; echo "default bench N=$N exit $?"
tail -n 1 gpurun_out/bench_n$N.log | python tools/brief.py | head -6; grep -i "graph capture failed\|Error" gpurun_out/bench_n$N.err | head -5
timeout 300 python -m pytest tests -m gpu -q -k "shard or gloo or dataparallel" 2>&1 | tail -2
