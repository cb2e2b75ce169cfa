#!/bin/bash
# run 40: full GPU suite + smoke at the last commit of the round
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r40_pytest.log 2>&1
grep -E "passed|failed|^E  " gpurun_out/r40_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r40_smoke.log 2>&1; tail -1 gpurun_out/r40_smoke.log
