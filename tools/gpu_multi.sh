#!/bin/bash
# N-GPU validation (N = $1): the default bench line (weak-scaling panoramas + the train workload's object, data-parallel
# with CUDA graphs around the NCCL gradient all-reduce) and the lattice-sharded 768x1536 workload.
N=${1:-2}
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_n$N.log 2>gpurun_out/bench_n$N.err; echo "default bench N=$N exit $?"; tail -n 1 gpurun_out/bench_n$N.log | cut -c1-300; grep -i "graph capture failed\|Error" gpurun_out/bench_n$N.err | head -5
timeout 600 $TR bench.py --gpus $N --workload pano768 --batch 8 --no-cpu-baseline > gpurun_out/bench_768_n$N.log 2>&1; echo "pano768 N=$N exit $?"; tail -n 1 gpurun_out/bench_768_n$N.log | cut -c1-300
timeout 600 python bench.py --workload pano768 --batch 8 --no-cpu-baseline > gpurun_out/bench_768_n1.log 2>&1; echo "pano768 N=1 exit $?"; tail -n 1 gpurun_out/bench_768_n1.log | cut -c1-300
