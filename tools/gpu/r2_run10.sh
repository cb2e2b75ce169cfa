#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hotpath.py tests/test_mode2d.py tests/test_interface_shim.py -m gpu -q --tb=short -p no:hypothesispytest > gpurun_out/r2_10_pytest.log 2>&1
grep -E "passed|failed|^E  |^FAILED" gpurun_out/r2_10_pytest.log | cut -c1-500
timeout 900 python bench.py --mode 2d > gpurun_out/r2_10_bench2d.log 2> gpurun_out/r2_10_bench2d.err
tail -c 2500 gpurun_out/r2_10_bench2d.log; tail -5 gpurun_out/r2_10_bench2d.err | cut -c1-400
timeout 600 python bench.py --mode 2d --impl reference --steps 2 --warmup 1 > gpurun_out/r2_10_bench2d_ref.log 2>&1
tail -c 700 gpurun_out/r2_10_bench2d_ref.log
