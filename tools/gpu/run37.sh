#!/bin/bash
# run 37: BASELINE config 1 shape (1k particles, box 128, 200 orientation samples = 25 x 8 phases) on one GPU, with the
# reference's CPU classes on the same shape beside it
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python bench.py --box 128 --particles 1000 --batch 1000 --mlr 25 --phases 8 --steps 4 --warmup 3 --cpu-sample 64 > gpurun_out/r37_bench_config1.log 2> gpurun_out/r37_bench_config1.err
tail -c 2600 gpurun_out/r37_bench_config1.log; tail -3 gpurun_out/r37_bench_config1.err
