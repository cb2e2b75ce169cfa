#!/bin/bash
# round 2, run 2: slab-ordered insert: parity, then bench A/B over the slab budget
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hotpath.py -m gpu -q --tb=short -x > gpurun_out/r2_02_pytest.log 2>&1
grep -E "passed|failed|^E  " gpurun_out/r2_02_pytest.log | cut -c1-300
for mb in 32 48 64; do
  THB_INSERT_SLAB_MB=$mb timeout 600 python bench.py --particles 10000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_02_bench_slab$mb.log 2> gpurun_out/r2_02_bench_slab$mb.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/r2_02_bench_slab$mb.log").read().strip().splitlines()[-1])
    print("slab_mb $mb value", round(j["value"],1), "ms/step", round(j["ms_per_step"],1), "insert", j["roofline"].get("insert_kernel"), "shares", j["roofline"]["share_of_step"], "clk", j["clocks"]["sm_mhz"])
except Exception as e:
    print("slab_mb $mb failed", e)
PY
done
THB_INSERT_IMPL=1 timeout 600 python bench.py --particles 10000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_02_bench_legacy.log 2> gpurun_out/r2_02_bench_legacy.err
tail -c 600 gpurun_out/r2_02_bench_legacy.log
