#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_iteration.py -m gpu -q --tb=short -s -p no:hypothesispytest > gpurun_out/r2_06_pytest.log 2>&1
grep -E "passed|failed|^E  |replay:|slot [01]:|2000 particles|^FAILED|iteration [0-9]:" gpurun_out/r2_06_pytest.log | cut -c1-900
timeout 300 python tools/gpu/pf_diag.py 48 > gpurun_out/r2_06_pf_diag.log 2>&1
grep -E "^[0-9] " gpurun_out/r2_06_pf_diag.log | cut -c1-300
