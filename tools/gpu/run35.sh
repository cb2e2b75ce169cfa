#!/bin/bash
# run 35: final validation of the round: full GPU suite, smoke, default bench, reference arm, ncu of the MODE_2D scan kernel
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r35_pytest.log 2>&1
grep -E "passed|failed|^E  " gpurun_out/r35_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r35_smoke.log 2>&1; tail -2 gpurun_out/r35_smoke.log
python bench.py > gpurun_out/r35_bench.log 2> gpurun_out/r35_bench.err; tail -c 2500 gpurun_out/r35_bench.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r35_bench_reference.log 2> gpurun_out/r35_bench_reference.err; tail -c 600 gpurun_out/r35_bench_reference.log
ncu --set full --clock-control none --import-source on -k regex:expect_direct -s 2 -c 1 -o gpurun_out/r35_prof_2dscan python tools/kbench2d.py --images 296 --reps 1 > gpurun_out/r35_ncu2d.log 2>&1
tail -3 gpurun_out/r35_ncu2d.log | cut -c1-300
