#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_mode2d.py -m gpu -q --tb=short -p no:hypothesispytest -k "all_classes" > gpurun_out/r2_30_pytest.log 2>&1
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_30_pytest.log | cut -c1-300 | head
