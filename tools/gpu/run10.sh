set -x
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/r10_clocks.csv &
SMI=$!
timeout 1500 python bench.py > gpurun_out/r10_bench.log 2> gpurun_out/r10_bench.err
kill $SMI
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r10_bench_ref.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r10_launches.csv python bench.py --particles 2000 --batch 1000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r10_launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:expect_direct -s 20 -c 1 -o gpurun_out/r10_prof_E python bench.py --particles 2000 --batch 1000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r10_ncuE.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:insert_kernel -s 1 -c 1 -o gpurun_out/r10_prof_M python bench.py --particles 2000 --batch 1000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r10_ncuM.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r10_pytest.log
tail -3 gpurun_out/r10_bench.log gpurun_out/r10_bench_ref.log gpurun_out/r10_pytest.log; tail -5 gpurun_out/r10_bench.err
