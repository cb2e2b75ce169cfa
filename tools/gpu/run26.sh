#!/bin/bash
# run 26: MODE_2D through both E kernels + the shims, 2D micro-benchmark, 3D bench regression check
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_mode2d.py tests/test_interface_shim.py -m gpu -q --tb=short > gpurun_out/r26_pytest_2d.log 2>&1
grep -E "passed|failed|^E  " gpurun_out/r26_pytest_2d.log | cut -c1-400
python tools/kbench2d.py --impl 3 > gpurun_out/r26_kbench2d_impl3.log 2>&1; cat gpurun_out/r26_kbench2d_impl3.log | cut -c1-400
python tools/kbench2d.py --impl 1 > gpurun_out/r26_kbench2d_impl1.log 2>&1; cat gpurun_out/r26_kbench2d_impl1.log | cut -c1-400
python bench.py --no-cpu-baseline > gpurun_out/r26_bench.log 2> gpurun_out/r26_bench.err
tail -c 1500 gpurun_out/r26_bench.log
