#!/bin/bash
# N-GPU validation (N = $1): weak-scaling panorama bench, data-parallel train bench, lattice-sharded 768x1536 bench.
N=${1:-2}
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1; echo "panorama N=$N exit $?"; tail -n 1 gpurun_out/bench_n$N.log | cut -c1-400
timeout 600 $TR bench.py --gpus $N --workload train --no-cpu-baseline > gpurun_out/bench_train_n$N.log 2>&1; echo "train N=$N exit $?"; tail -n 1 gpurun_out/bench_train_n$N.log | cut -c1-400
timeout 600 $TR bench.py --gpus $N --workload pano768 --batch 8 --no-cpu-baseline > gpurun_out/bench_768_n$N.log 2>&1; echo "pano768 N=$N exit $?"; tail -n 1 gpurun_out/bench_768_n$N.log | cut -c1-400
timeout 600 python bench.py --workload pano768 --batch 8 --no-cpu-baseline > gpurun_out/bench_768_n1.log 2>&1; echo "pano768 N=1 exit $?"; tail -n 1 gpurun_out/bench_768_n1.log | cut -c1-400
