#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 ./tools/gpu/bin/tmabench 2>&1 | tee gpurun_out/r2_26_tmabench.log
