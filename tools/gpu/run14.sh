set -x
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r14_pytest.log
for k in 1e-6 1.5e-5 1e-3; do
 for b in 0 2 3; do
  for s in 0 1; do
   THB_QUAD_BRICK=$b THB_SORT_ROT=$s python tools/kbench.py 1024 256 $k 2>&1 | grep "^E:" | tail -1 | sed "s/^/k=$k brick=$b sort=$s /" >> gpurun_out/r14_sweep.log
  done
 done
done
cat gpurun_out/r14_sweep.log; tail -5 gpurun_out/r14_pytest.log
