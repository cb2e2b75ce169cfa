#!/bin/bash
# one arrival counter per barrier in the lockstep launch (exact window): the tests that exercise it and a short bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_iteration.py -m gpu -q --tb=short -p no:hypothesispytest -k "lockstep or kernels_agree or adaptive or box256 or fsc_gate or replay" > gpurun_out/r2_28_pytest.log 2>&1
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_28_pytest.log | cut -c1-300 | head
for w in 1 0; do
THB_EXPECT_LOCK_WINDOW=$w timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_28_bench_w$w.log 2> gpurun_out/r2_28_bench_w$w.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_28_bench_w$w.log") if l.startswith("{")][-1])
    print("window $w: value", round(j["value"],1), "ms/step", round(j["ms_per_step"],1), "E ms", round(j["roofline"]["avg_launch_ms"],1), "clk", j["clocks"]["sm_mhz"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_28_bench_w$w.err").read()[-1500:])
PY
done
