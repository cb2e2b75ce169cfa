#!/bin/bash
# run 23: sigma-refresh parity, closed-loop iterations, default bench (oct layout)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_reco_oracle.py tests/test_gpu_iteration.py -m gpu -q -s --tb=short -k "sigma or closed_loop" > gpurun_out/r23_pytest.log 2>&1
grep -E "iteration|passed|failed|^E  " gpurun_out/r23_pytest.log | cut -c1-400
python bench.py > gpurun_out/r23_bench.log 2> gpurun_out/r23_bench.err
tail -c 3000 gpurun_out/r23_bench.log
