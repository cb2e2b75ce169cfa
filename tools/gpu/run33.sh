#!/bin/bash
# run 33: MODE_2D reconstruct / set_projectee + regression of the 3D reconstruction tests and the shims
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_mode2d.py tests/test_reco_oracle.py tests/test_interface_shim.py -m gpu -q --tb=short > gpurun_out/r33_pytest.log 2>&1
grep -E "passed|failed|^E  " gpurun_out/r33_pytest.log | cut -c1-400
