#!/bin/bash
# run 31: pixels-on-lanes E kernel (expect_impl 5): parity, micro-benchmark against impl 3, bench
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_hotpath.py tests/test_mode2d.py -m gpu -q --tb=short -k "kernels_agree or symmetrize or 2d" > gpurun_out/r31_pytest.log 2>&1
grep -E "passed|failed|^E  " gpurun_out/r31_pytest.log | cut -c1-300
for k in 1e-6 1.5e-5 1e-3; do
  for impl in 3 5; do
    THB_EXPECT_IMPL=$impl python tools/kbench.py 1024 256 $k 2>&1 | grep -E "^E:" | tail -1 | sed "s/^/k=$k impl=$impl /" >> gpurun_out/r31_kbench.log
  done
done
cat gpurun_out/r31_kbench.log
THB_EXPECT_IMPL=5 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r31_bench_impl5.log 2> gpurun_out/r31_bench_impl5.err
tail -c 1200 gpurun_out/r31_bench_impl5.log
THB_EXPECT_IMPL=3 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r31_bench_impl3.log 2> gpurun_out/r31_bench_impl3.err
tail -c 1200 gpurun_out/r31_bench_impl3.log
