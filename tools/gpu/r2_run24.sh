#!/bin/bash
# the adaptive-compaction test again, and compute-sanitizer (memcheck) over the kernels added this round
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_iteration.py -m gpu -q --tb=short -k "adaptive or global_search" -p no:hypothesispytest > gpurun_out/r2_24_pytest.log 2>&1
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_24_pytest.log | cut -c1-300 | head
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hotpath.py tests/test_mode2d.py tests/test_gpu_iteration.py -m gpu -q -x --tb=line -p no:hypothesispytest \
  -k "scan_matches_local or all_classes_in_one_launch or kernels_agree_with_each_other_and_oracle[2e-05-125-9] or adaptive or device_pf_operators" > gpurun_out/r2_24_memcheck.log 2>&1
echo "memcheck exit $?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" gpurun_out/r2_24_memcheck.log | sort | uniq -c | head
