set -x
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r11_pytest.log
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r11_bench.log 2> gpurun_out/r11_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r11_launches.csv python bench.py --particles 2000 --batch 1000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r11_launches_run.log 2>&1
tail -3 gpurun_out/r11_pytest.log gpurun_out/r11_bench.log; tail -3 gpurun_out/r11_bench.err
