#!/bin/bash
# run 38: 15-translations-per-pass variant of the default E kernel: parity (multi-translation shapes, 2D scans) and the 2D scan rate
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_hotpath.py tests/test_mode2d.py tests/test_interface_shim.py -m gpu -q --tb=short > gpurun_out/r38_pytest.log 2>&1
grep -E "passed|failed|^E  " gpurun_out/r38_pytest.log | cut -c1-300
python tools/kbench2d.py --impl 3 > gpurun_out/r38_kbench2d.log 2>&1; cat gpurun_out/r38_kbench2d.log | cut -c1-400
