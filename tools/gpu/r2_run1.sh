#!/bin/bash
# round 2, run 1: L2-footprint micro-benchmark (gathers and reductions)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_01_l2bench.log
timeout 300 ./tools/gpu/bin/l2bench >> gpurun_out/r2_01_l2bench.log 2>&1
cat gpurun_out/r2_01_l2bench.log
