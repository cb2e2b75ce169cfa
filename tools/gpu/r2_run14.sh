#!/bin/bash
# several rotations per lane (expect_impl 7): parity, kernel micro-benchmark over cloud widths, bench A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hotpath.py -m gpu -q --tb=short -k "kernels_agree or expect_local" -p no:hypothesispytest > gpurun_out/r2_14_pytest.log 2>&1
grep -E "passed|failed|^E  |^FAILED" gpurun_out/r2_14_pytest.log | cut -c1-400
for k in 1e-6 1.5e-5 1e-3; do
  KBENCH_NO_INSERT=1 KBENCH_IMPLS=3,7:2,7:4 timeout 600 python tools/kbench.py 1024 256 $k 2>&1 | grep "^E\[" | awk 'NR%4==0' | tee -a gpurun_out/r2_14_kbench.log
done
for v in "3 2" "7 2" "7 4"; do
  set -- $v
  THB_EXPECT_IMPL=$1 THB_EXPECT_RPL=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_14_bench_impl$1_rpl$2.log 2> gpurun_out/r2_14_bench_impl$1_rpl$2.err
  python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_14_bench_impl$1_rpl$2.log") if l.startswith("{")][-1])
    print("impl $1 rpl $2: value", round(j["value"],1), "ms/step", round(j["ms_per_step"],1), "frac", round(j["roofline"]["frac"],3), "E ms", round(j["roofline"]["avg_launch_ms"],1), "shares", {k: round(v,3) for k,v in j["roofline"]["share_of_step"].items()}, "clk", j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_14_bench_impl$1_rpl$2.err").read()[-1500:])
PY
done
