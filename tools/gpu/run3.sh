set -x
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r3_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_iteration.py -m gpu -x -q -k "staged or golden or phase_equals" > gpurun_out/r3_sanitizer.log 2>&1; echo "sanitizer rc $?" >> gpurun_out/r3_sanitizer.log
for k in 1e-6 7.6e-5 1e-3; do
  python tools/kbench.py 1024 256 $k > gpurun_out/r3_kbench_v2_$k.log 2>&1
done
THB_INSERT_IMPL=1 THB_EXPECT_IMPL=1 python tools/kbench.py 1024 256 7.6e-5 > gpurun_out/r3_kbench_v1_ins1_7.6e-5.log 2>&1
timeout 900 python bench.py --particles 20000 --batch 2000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3_bench_small.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:expect_local_tma -s 1 -c 1 -o gpurun_out/r3_prof_E python tools/kbench.py 296 256 7.6e-5 > gpurun_out/r3_ncuE.log 2>&1
for f in gpurun_out/r3_*.log; do echo "== $f"; tail -n 8 $f; done
