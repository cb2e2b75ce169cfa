set -x
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r4_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_iteration.py -m gpu -x -q -k "staged or golden or phase_equals" > gpurun_out/r4_sanitizer.log 2>&1; echo "sanitizer rc $?" >> gpurun_out/r4_sanitizer.log
for k in 1e-6 1.5e-5 7.6e-5; do
  python tools/kbench.py 1024 256 $k > gpurun_out/r4_kbench_t16x8_$k.log 2>&1
  THB_TILE_W=8 THB_TILE_H=8 python tools/kbench.py 1024 256 $k > gpurun_out/r4_kbench_t8x8_$k.log 2>&1
done
THB_EXPECT_IMPL=1 python tools/kbench.py 1024 256 1.5e-5 > gpurun_out/r4_kbench_v1_1.5e-5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:expect_local_tma -s 1 -c 1 -o gpurun_out/r4_prof_E python tools/kbench.py 296 256 1.5e-5 > gpurun_out/r4_ncuE.log 2>&1
for f in gpurun_out/r4_*.log; do echo "== $f"; tail -n 8 $f | grep -E "^E:|passed|failed|rc|Error|error" | tail -3; done
