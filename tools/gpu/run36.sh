#!/bin/bash
# run 36: one GPU's share of BASELINE config 4 (20k particles, box 512, 4000 orientation samples, 4 GPUs -> 5000 particles/GPU)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 700 python bench.py --box 512 --particles 5000 --batch 500 --phases 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r36_bench_box512.log 2> gpurun_out/r36_bench_box512.err
tail -c 2500 gpurun_out/r36_bench_box512.log; tail -5 gpurun_out/r36_bench_box512.err
