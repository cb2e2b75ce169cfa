#!/bin/bash
# run 32 (2 GPUs): NCCL tests + the driver's own launch line of the bench at N = 2, default sizes
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r32_pytest_multi.log
tail -4 gpurun_out/r32_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r32_bench_2gpu.log 2> gpurun_out/r32_bench_2gpu.err
tail -2 gpurun_out/r32_bench_2gpu.log | cut -c1-1500; tail -3 gpurun_out/r32_bench_2gpu.err
