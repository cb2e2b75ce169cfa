#!/bin/bash
# run 25: MODE_2D parity tests + the rest of the GPU suite (THB_MAX_SLOTS change touches every kernel's argument block)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_mode2d.py -m gpu -q --tb=short > gpurun_out/r25_pytest_2d.log 2>&1
grep -E "passed|failed|^E  " gpurun_out/r25_pytest_2d.log | cut -c1-400
python -m pytest tests -m gpu -q --tb=short --deselect tests/test_mode2d.py > gpurun_out/r25_pytest.log 2>&1
grep -E "passed|failed|^E  " gpurun_out/r25_pytest.log | cut -c1-400
