set -x
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r5_pytest.log
for k in 1e-6 1.5e-5 7.6e-5; do
  python tools/kbench.py 1024 256 $k > gpurun_out/r5_kbench_t16x8_$k.log 2>&1
  THB_TILE_W=8 THB_TILE_H=8 python tools/kbench.py 1024 256 $k > gpurun_out/r5_kbench_t8x8_$k.log 2>&1
done
THB_INSERT_IMPL=2 python tools/kbench.py 512 256 1.5e-5 > gpurun_out/r5_kbench_ins2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:expect_local_tma -s 1 -c 1 -o gpurun_out/r5_prof_E python tools/kbench.py 296 256 1.5e-5 > gpurun_out/r5_ncuE.log 2>&1
for f in gpurun_out/r5_*.log; do echo "== $f"; grep -E "^E:|^M:|staging|passed|failed" $f | tail -4; done
