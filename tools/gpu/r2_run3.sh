#!/bin/bash
# round 2, run 3: replay parity chain + FSC gate, hot-path suite, default bench, ncu of the slab insert
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_iteration.py tests/test_gpu_hotpath.py -m gpu -q --tb=short -s -p no:hypothesispytest > gpurun_out/r2_03_pytest.log 2>&1
grep -E "passed|failed|^E  |replay:|slot [01]:|best orientation|^FAILED" gpurun_out/r2_03_pytest.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/r2_03_bench.log 2> gpurun_out/r2_03_bench.err
tail -c 2500 gpurun_out/r2_03_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:insert_slab -c 1 -o gpurun_out/r2_03_insert_slab python bench.py --particles 5000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_03_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
