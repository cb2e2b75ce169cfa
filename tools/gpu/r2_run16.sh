#!/bin/bash
# new defaults (lockstep E kernel on the radial order, shared-template scan): full GPU suite, ncu of the E kernel, sweeps, 2D
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:hypothesispytest > gpurun_out/r2_16_pytest.log 2>&1
grep -E "passed|failed|^E  |^FAILED" gpurun_out/r2_16_pytest.log | cut -c1-300 | head -30
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_16_bench_$name.log 2> gpurun_out/r2_16_bench_$name.err
  python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_16_bench_$name.log") if l.startswith("{")][-1])
    print("$name: value", round(j["value"],1), "ms/step", round(j["ms_per_step"],1), "frac", round(j["roofline"]["frac"],3), "E ms", round(j["roofline"]["avg_launch_ms"],1), "clk", j["clocks"]["sm_mhz"], {k: round(v,3) for k,v in j["roofline"]["share_of_step"].items()})
except Exception as e:
    print("$name: bench failed", e); print(open("gpurun_out/r2_16_bench_$name.err").read()[-1200:])
PY
}
run default
run tiles2 THB_EXPECT_LOCK_TILES=2
run tiles4_w0 THB_EXPECT_LOCK_TILES=4 THB_EXPECT_LOCK_WINDOW=0
run rpl4 THB_EXPECT_RPL=4
for st in 0 1; do
  THB_SCAN_TEMPLATES=$st timeout 600 python tools/kbench2d.py --images 2048 2>&1 | grep "^scan" | sed "s/^/templates=$st /" | tee -a gpurun_out/r2_16_kbench2d.log
done
timeout 900 python bench.py --mode 2d --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_16_bench_2d.log 2> gpurun_out/r2_16_bench_2d.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_16_bench_2d.log") if l.startswith("{")][-1])
    print("2d: value", round(j["value"],1), "e2e", j["e2e"] and round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],1))
except Exception as e:
    print("2d bench failed", e); print(open("gpurun_out/r2_16_bench_2d.err").read()[-1500:])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:expect_multi -s 20 -c 1 -o gpurun_out/r2_16_ncu_expect_multi python bench.py --particles 10000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_16_ncu.log 2>&1
ncu -i gpurun_out/r2_16_ncu_expect_multi.ncu-rep --page details > gpurun_out/r2_16_ncu_expect_multi_details.txt 2>&1
grep -E "Duration|DRAM Throughput|L1/TEX Hit|L2 Hit|L1/TEX Cache Throughput|L2 Cache Throughput|Issue Slots Busy|Registers Per|dram__bytes_read" gpurun_out/r2_16_ncu_expect_multi_details.txt | head -12
ls -la gpurun_out/*.ncu-rep
