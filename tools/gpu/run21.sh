set -x
THB_QUAD_OCT=1 python -m pytest tests/test_gpu_hotpath.py -m gpu -q --tb=line -k "kernels_agree or golden or box256" 2>&1 | tail -4 > gpurun_out/r21_pytest_oct.log
for k in 1e-6 1.5e-5 1e-3; do
  for o in 0 1; do
   THB_QUAD_OCT=$o python tools/kbench.py 1024 256 $k 2>&1 | grep "^E:" | tail -1 | sed "s/^/k=$k oct=$o /" >> gpurun_out/r21_sweep.log
  done
done
THB_QUAD_OCT=1 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r21_bench_oct.log 2>&1
cat gpurun_out/r21_sweep.log; tail -3 gpurun_out/r21_pytest_oct.log; tail -1 gpurun_out/r21_bench_oct.log | cut -c1-200
