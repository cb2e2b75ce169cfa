#!/bin/bash
# MODE_2D after the round's changes: tests, bench.py --mode 2d (whole iteration on the device) with its reference arm, ncu of the scan contraction
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mode2d.py tests/test_gpu_hotpath.py tests/test_interface_shim.py -m gpu -q --tb=short -k "2d or 2D or scan or classification" -p no:hypothesispytest > gpurun_out/r2_19_pytest.log 2>&1
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_19_pytest.log | cut -c1-300 | head
timeout 900 python bench.py --mode 2d > gpurun_out/r2_19_bench_2d.log 2> gpurun_out/r2_19_bench_2d.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_19_bench_2d.log") if l.startswith("{")][-1])
    print("2d: value", round(j["value"],1), "e2e", j["e2e"] and round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],1), "shares", {k: round(v,3) for k,v in j["roofline"]["share_of_step"].items()}, "T terms/s", j["roofline"]["pixel_rot_trans_per_s"], "cpu", j.get("cpu_baseline"), "clk", j["clocks"])
except Exception as e:
    print("2d bench failed", e); print(open("gpurun_out/r2_19_bench_2d.err").read()[-2500:])
PY
timeout 900 python bench.py --mode 2d --impl reference --steps 2 --warmup 1 > gpurun_out/r2_19_bench_2d_reference.log 2> gpurun_out/r2_19_bench_2d_reference.err
tail -c 600 gpurun_out/r2_19_bench_2d_reference.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_contract -s 3 -c 1 -o gpurun_out/r2_19_ncu_scan_contract python tools/kbench2d.py --images 592 --reps 1 > gpurun_out/r2_19_ncu.log 2>&1
ncu -i gpurun_out/r2_19_ncu_scan_contract.ncu-rep --page details > gpurun_out/r2_19_ncu_scan_contract_details.txt 2>&1
grep -E "Duration|DRAM Throughput|L1/TEX Hit|L2 Hit|L1/TEX Cache Throughput|L2 Cache Throughput|Issue Slots Busy|Registers Per|Achieved Occupancy|FMA is|highest-utilized|Stall|stalled" gpurun_out/r2_19_ncu_scan_contract_details.txt | cut -c1-220 | head -16
