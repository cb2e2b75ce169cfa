#!/bin/bash
# run 24: full GPU suite (sigma refresh, async upload, closed loop) + default bench with overlapped e2e uploads
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s --tb=short > gpurun_out/r24_pytest.log 2>&1
grep -E "^iteration|passed|failed|^E  " gpurun_out/r24_pytest.log | cut -c1-400
python bench.py > gpurun_out/r24_bench.log 2> gpurun_out/r24_bench.err
tail -c 3000 gpurun_out/r24_bench.log; tail -5 gpurun_out/r24_bench.err
