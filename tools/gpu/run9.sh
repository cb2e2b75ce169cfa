set -x
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r9_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hotpath.py -m gpu -x -q -k "kernels_agree or golden" > gpurun_out/r9_sanitizer.log 2>&1; echo "sanitizer rc $?" >> gpurun_out/r9_sanitizer.log
for k in 1e-6 4e-6 1.5e-5 7.6e-5 1e-3; do
  THB_EXPECT_MINB=2 python tools/kbench.py 1024 256 $k > gpurun_out/r9_kbench_q2_$k.log 2>&1
  THB_EXPECT_MINB=3 python tools/kbench.py 1024 256 $k > gpurun_out/r9_kbench_q3_$k.log 2>&1
done
THB_EXPECT_IMPL=1 python tools/kbench.py 1024 256 4e-6 > gpurun_out/r9_kbench_v1_4e-6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:expect_direct -s 1 -c 1 -o gpurun_out/r9_prof_E python tools/kbench.py 296 256 1.5e-5 > gpurun_out/r9_ncuE.log 2>&1
for f in gpurun_out/r9_*.log; do echo "== $f"; grep -E "^E:|passed|failed|rc" $f | sed -n '1p;$p'; done
