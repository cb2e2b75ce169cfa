#!/bin/bash
# run 34 (8 GPUs): the driver's own launch line of the bench at N = 8, default sizes, short
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r34_bench_8gpu.log 2> gpurun_out/r34_bench_8gpu.err
tail -1 gpurun_out/r34_bench_8gpu.log | cut -c1-1600; tail -3 gpurun_out/r34_bench_8gpu.err
