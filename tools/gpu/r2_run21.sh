#!/bin/bash
# all-classes 2D scan (thb_expect_scan_classes): parity, bench.py --mode 2d; TMA-vs-LSU gather micro-benchmark
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mode2d.py tests/test_interface_shim.py -m gpu -q --tb=short -p no:hypothesispytest > gpurun_out/r2_21_pytest.log 2>&1
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_21_pytest.log | cut -c1-300 | head
timeout 900 python bench.py --mode 2d > gpurun_out/r2_21_bench_2d.log 2> gpurun_out/r2_21_bench_2d.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_21_bench_2d.log") if l.startswith("{")][-1])
    print("2d: value", round(j["value"],1), "e2e", j["e2e"] and round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],1), "shares", {k: round(v,3) for k,v in j["roofline"]["share_of_step"].items()}, "T terms/s", j["roofline"]["pixel_rot_trans_per_s"], "cpu", j.get("cpu_baseline"), "clk", j["clocks"]["sm_mhz"])
except Exception as e:
    print("2d bench failed", e); print(open("gpurun_out/r2_21_bench_2d.err").read()[-2500:])
PY
timeout 300 ./tools/gpu/bin/tmabench 2>&1 | tee gpurun_out/r2_21_tmabench.log
