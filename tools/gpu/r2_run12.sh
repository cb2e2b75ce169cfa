#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mode2d.py -k classification_iteration -m gpu -q --tb=short -s -p no:hypothesispytest > gpurun_out/r2_12_pytest.log 2>&1
grep -E "passed|failed|^E  |^FAILED|2D iteration|after the" gpurun_out/r2_12_pytest.log | cut -c1-600
