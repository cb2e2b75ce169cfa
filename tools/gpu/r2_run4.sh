#!/bin/bash
# round 2, run 4: device-vs-host PF operators with the same stream, replay chain diagnostics, reference-protocol seam, spread kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_iteration.py tests/test_interface_shim.py -m gpu -q --tb=short -s -p no:hypothesispytest > gpurun_out/r2_04_pytest.log 2>&1
grep -E "passed|failed|^E  |replay:|slot [01]:|2000 particles|^FAILED|iteration [0-9]:" gpurun_out/r2_04_pytest.log | cut -c1-600
timeout 900 python -m pytest tests/test_gpu_hotpath.py tests/test_mode2d.py tests/test_reco_oracle.py -m gpu -q --tb=short -p no:hypothesispytest > gpurun_out/r2_04_pytest_b.log 2>&1
grep -E "passed|failed|^E  |^FAILED" gpurun_out/r2_04_pytest_b.log | cut -c1-400
