#!/bin/bash
# Run on the GPU box via gpurun: every group in its own process so that a trapped kernel cannot poison the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py build > gpurun_out/build.log 2>&1
run() { # name, timeout, command...
  local name=$1; local t=$2; shift 2
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "$name exit $?" | tee -a gpurun_out/summary.txt
  tail -n 12 gpurun_out/$name.log
}
: > gpurun_out/summary.txt
run ops_basic 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "bias_act or upfirdn or gather or linear or noise"
run conv_simt 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "simt or module or rgb"
run conv_tcgen05 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "tcgen05"
run generator 900 python -m pytest tests/test_gpu_generator.py -m gpu -q
run smoke 300 python __graft_entry__.py smoke
run bench_small 900 python bench.py --batch 8 --steps 1 --warmup 3 --cpu-sample-patches 6
