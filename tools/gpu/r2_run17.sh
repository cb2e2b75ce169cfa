#!/bin/bash
# validation of the round's defaults: full GPU suite, the default bench line (with e2e and CPU baseline), ncu launch list,
# ncu of the particle-filter kernel, BASELINE config 1 shape
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:hypothesispytest > gpurun_out/r2_17_pytest.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r2_17_pytest.log | cut -c1-300 | head -30
timeout 900 python bench.py > gpurun_out/r2_17_bench_default.log 2> gpurun_out/r2_17_bench_default.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_17_bench_default.log") if l.startswith("{")][-1])
    print("default: value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],1), "frac", round(j["roofline"]["frac"],3), "E ms", round(j["roofline"]["avg_launch_ms"],1), "clk", j["clocks"], {k: round(v,3) for k,v in j["roofline"]["share_of_step"].items()}, "cpu", j["cpu_baseline"], "launches", j["gpu_launches"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_17_bench_default.err").read()[-1500:])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_17_bench_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_17_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pf_step -s 6 -c 1 -o gpurun_out/r2_17_ncu_pf_step python bench.py --particles 5000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_17_ncu_pf.log 2>&1
ncu -i gpurun_out/r2_17_ncu_pf_step.ncu-rep --page details > gpurun_out/r2_17_ncu_pf_step_details.txt 2>&1
grep -E "Duration|Issue Slots Busy|Registers Per|Achieved Occupancy|Theoretical Occupancy|Executed Ipc Active" gpurun_out/r2_17_ncu_pf_step_details.txt | head
timeout 600 python bench.py --box 128 --particles 1000 --batch 1000 --mlr 25 --phases 8 --cpu-sample 64 > gpurun_out/r2_17_bench_config1.log 2> gpurun_out/r2_17_bench_config1.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_17_bench_config1.log") if l.startswith("{")][-1])
    print("config1: value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],2), "frac", round(j["roofline"]["frac"],3), "cpu", j["cpu_baseline"]["value"])
except Exception as e:
    print("config1 bench failed", e); print(open("gpurun_out/r2_17_bench_config1.err").read()[-1500:])
PY
