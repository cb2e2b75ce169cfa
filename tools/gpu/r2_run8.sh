#!/bin/bash
# round 2, run 8: ncu launch list of the bench command and ncu --set full of the dominant E kernel at the bench's own launch size
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_08_launches.csv python bench.py --particles 10000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_08_launches.log 2>&1
tail -3 gpurun_out/r2_08_launches.csv | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:expect_direct -s 20 -c 1 -o gpurun_out/r2_08_expect python bench.py --particles 10000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_08_ncuE.log 2>&1
ls -la gpurun_out/r2_08_expect.ncu-rep
