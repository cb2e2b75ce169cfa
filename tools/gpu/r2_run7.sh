#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ctf_search.py tests/test_interface_shim.py -m gpu -q --tb=short -s -p no:hypothesispytest > gpurun_out/r2_07_pytest.log 2>&1
grep -E "passed|failed|^E  |^FAILED" gpurun_out/r2_07_pytest.log | cut -c1-600
timeout 900 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_multi.py tests/test_mode2d.py -m gpu -q --tb=short -p no:hypothesispytest > gpurun_out/r2_07_pytest_b.log 2>&1
grep -E "passed|failed|^E  |^FAILED" gpurun_out/r2_07_pytest_b.log | cut -c1-400
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_07_bench.log 2> gpurun_out/r2_07_bench.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/r2_07_bench.log").read().strip().splitlines()[-1])
    print("value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "steps", j["e2e"]["steps"], "ms/step", round(j["ms_per_step"],1), "frac", round(j["roofline"]["frac"],3), "shares", {k: round(v,3) for k,v in j["roofline"]["share_of_step"].items()}, "clk", j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_07_bench.err").read()[-1500:])
PY
