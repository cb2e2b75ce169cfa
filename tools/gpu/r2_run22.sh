#!/bin/bash
# global-search iteration at the config-1 shape through bench.py (scan of 10 000 rotations x 30 translations + hand-over + phases + insert),
# the tests touched by the last changes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mode2d.py tests/test_interface_shim.py tests/test_gpu_iteration.py -m gpu -q --tb=short -s -k "scan or 2d or 2D or global or classification or handover" -p no:hypothesispytest > gpurun_out/r2_22_pytest.log 2>&1
grep -E "passed|failed|^FAILED|^E  |worst relative|global search:" gpurun_out/r2_22_pytest.log | cut -c1-250 | sort | uniq -c | sort -rn | head -30
timeout 900 python bench.py --box 128 --particles 1000 --batch 1000 --mlr 25 --phases 8 --scan-nr 10000 --nt 30 --cpu-sample 32 --steps 3 --warmup 2 > gpurun_out/r2_22_bench_global.log 2> gpurun_out/r2_22_bench_global.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_22_bench_global.log") if l.startswith("{")][-1])
    print("global: value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],2), "shares", {k: round(v,3) for k,v in j["roofline"]["share_of_step"].items()}, "cpu", j["cpu_baseline"], "launches", j["gpu_launches"])
except Exception as e:
    print("global bench failed", e); print(open("gpurun_out/r2_22_bench_global.err").read()[-2500:])
PY
