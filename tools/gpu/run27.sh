#!/bin/bash
# run 27: normCorrection parity + reco/sigma suite
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_reco_oracle.py -m gpu -q --tb=short > gpurun_out/r27_pytest.log 2>&1
grep -E "passed|failed|^E  " gpurun_out/r27_pytest.log | cut -c1-400
