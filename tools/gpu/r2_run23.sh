#!/bin/bash
# full GPU suite + the default bench line on the final tree of the round
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:hypothesispytest > gpurun_out/r2_23_pytest.log 2>&1
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_23_pytest.log | cut -c1-300 | head -30
timeout 900 python bench.py > gpurun_out/r2_23_bench_default.log 2> gpurun_out/r2_23_bench_default.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_23_bench_default.log") if l.startswith("{")][-1])
    print("default: value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],1), "frac", round(j["roofline"]["frac"],3), "E ms", round(j["roofline"]["avg_launch_ms"],1), "clk", j["clocks"], {k: round(v,3) for k,v in j["roofline"]["share_of_step"].items()}, "cpu", j["cpu_baseline"]["value"], "launches", j["gpu_launches"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_23_bench_default.err").read()[-1500:])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
