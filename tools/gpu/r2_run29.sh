#!/bin/bash
# the full GPU suite and smoke() on the final tree of the round
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --tb=short -p no:hypothesispytest > gpurun_out/r2_29_pytest.log 2>&1
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_29_pytest.log | cut -c1-300 | head -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
