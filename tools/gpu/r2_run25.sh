#!/bin/bash
# adaptive E-step (data-dependent phase counts) with and without the compaction of unfinished particles; ncu of the all-classes scan
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for c in 1 0; do
  THB_PF_COMPACT=$c timeout 900 python bench.py --phases 0 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2_25_bench_adaptive_compact$c.log 2> gpurun_out/r2_25_bench_adaptive_compact$c.err
  python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_25_bench_adaptive_compact$c.log") if l.startswith("{")][-1])
    print("adaptive, compaction $c: value", round(j["value"],1), "ms/step", round(j["ms_per_step"],1), "E launches per step", j["roofline"]["launches"] / j["steps"], "avg E ms", round(j["roofline"]["avg_launch_ms"],1), {k: round(v,3) for k,v in j["roofline"]["share_of_step"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_25_bench_adaptive_compact$c.err").read()[-1500:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_contract -s 1 -c 1 -o gpurun_out/r2_25_ncu_scan_classes python bench.py --mode 2d --particles 5000 --batch 1184 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_25_ncu.log 2>&1
ncu -i gpurun_out/r2_25_ncu_scan_classes.ncu-rep --page details > gpurun_out/r2_25_ncu_scan_classes_details.txt 2>&1
grep -E "Duration|L1/TEX Cache Throughput|L2 Cache Throughput|Issue Slots Busy|Registers Per|Achieved Occupancy|highest-utilized|L2 Hit|L1/TEX Hit|bank conflict" gpurun_out/r2_25_ncu_scan_classes_details.txt | cut -c1-200 | head -12
