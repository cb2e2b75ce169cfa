#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ctf_search.py tests/test_gpu_iteration.py -m gpu -q --tb=short -s -p no:hypothesispytest > gpurun_out/r2_09_pytest.log 2>&1
grep -E "passed|failed|^E  |^FAILED|CTF search:" gpurun_out/r2_09_pytest.log | cut -c1-600
