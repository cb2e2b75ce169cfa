set -x
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25 > gpurun_out/r22_pytest.log
grep -E "passed|failed|^E  " gpurun_out/r22_pytest.log | cut -c1-300
python -m pytest tests/test_gpu_iteration.py -m gpu -q -s -k closed_loop 2>&1 | grep -E "iteration|passed|failed" | cut -c1-300 > gpurun_out/r22_closed_loop.log; cat gpurun_out/r22_closed_loop.log
ncu --set full --clock-control none --import-source on -k regex:expect_direct -s 20 -c 1 -o gpurun_out/r22_prof_E python bench.py --particles 2000 --batch 1000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r22_ncuE.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r22_launches.csv python bench.py --particles 2000 --batch 1000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r22_launches_run.log 2>&1
