#!/bin/bash
# lockstep launch on the radial pixel order (L2-resident shell working set): parity, then bench sweeps; seam executable; pack CTF
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hotpath.py tests/test_interface_shim.py -m gpu -q --tb=short -k "kernels_agree or lockstep or pack_stack or insertI or local_search_through" -p no:hypothesispytest > gpurun_out/r2_15_pytest.log 2>&1
grep -E "passed|failed|^E  |^FAILED" gpurun_out/r2_15_pytest.log | cut -c1-400
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_15_bench_$name.log 2> gpurun_out/r2_15_bench_$name.err
  python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_15_bench_$name.log") if l.startswith("{")][-1])
    print("$name: value", round(j["value"],1), "ms/step", round(j["ms_per_step"],1), "frac", round(j["roofline"]["frac"],3), "E ms", round(j["roofline"]["avg_launch_ms"],1), "clk", j["clocks"]["sm_mhz"])
except Exception as e:
    print("$name: bench failed", e); print(open("gpurun_out/r2_15_bench_$name.err").read()[-1200:])
PY
}
run base_impl7 THB_EXPECT_IMPL=7
run radial_nolock THB_EXPECT_IMPL=7 THB_EXPECT_ORDER=1
run lock_oct_w2 THB_EXPECT_IMPL=7 THB_EXPECT_ORDER=1 THB_EXPECT_LOCK=1 THB_EXPECT_LOCK_WINDOW=2
run lock_quad_w2 THB_EXPECT_IMPL=7 THB_EXPECT_ORDER=1 THB_EXPECT_LOCK=1 THB_EXPECT_LOCK_WINDOW=2 THB_QUAD_OCT=0
run lock_quad_w0 THB_EXPECT_IMPL=7 THB_EXPECT_ORDER=1 THB_EXPECT_LOCK=1 THB_EXPECT_LOCK_WINDOW=0 THB_QUAD_OCT=0
run lock_quad_w6 THB_EXPECT_IMPL=7 THB_EXPECT_ORDER=1 THB_EXPECT_LOCK=1 THB_EXPECT_LOCK_WINDOW=6 THB_QUAD_OCT=0
run lock_quad_w2_blocked THB_EXPECT_IMPL=7 THB_EXPECT_ORDER=0 THB_EXPECT_LOCK=1 THB_EXPECT_LOCK_WINDOW=2 THB_QUAD_OCT=0
