#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/gpu/pf_diag.py 48 > gpurun_out/r2_05_pf_diag.log 2>&1
cat gpurun_out/r2_05_pf_diag.log | cut -c1-900
