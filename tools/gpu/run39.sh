#!/bin/bash
# run 39: adaptive rotation/pixel split: parity, headline bench regression check, config-1 shape, 2D scan
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_hotpath.py tests/test_mode2d.py tests/test_gpu_iteration.py -m gpu -q --tb=short > gpurun_out/r39_pytest.log 2>&1
grep -E "passed|failed|^E  " gpurun_out/r39_pytest.log | cut -c1-300
python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r39_bench.log 2> gpurun_out/r39_bench.err; tail -c 900 gpurun_out/r39_bench.log
python bench.py --box 128 --particles 1000 --batch 1000 --mlr 25 --phases 8 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r39_bench_config1.log 2> gpurun_out/r39_bench_config1.err; tail -c 700 gpurun_out/r39_bench_config1.log
python tools/kbench2d.py --impl 3 > gpurun_out/r39_kbench2d.log 2>&1; tail -2 gpurun_out/r39_kbench2d.log | cut -c1-300
