set -x
./tools/gpu/bin/gatherbench > gpurun_out/r7_gatherbench.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r7_pytest.log
for k in 1e-6 1.5e-5 7.6e-5; do
  python tools/kbench.py 1024 256 $k > gpurun_out/r7_kbench_t16x8_$k.log 2>&1
  THB_TILE_W=8 THB_TILE_H=8 python tools/kbench.py 1024 256 $k > gpurun_out/r7_kbench_t8x8_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:expect_local_tma -s 1 -c 1 -o gpurun_out/r7_prof_E python tools/kbench.py 296 256 1.5e-5 > gpurun_out/r7_ncuE.log 2>&1
cat gpurun_out/r7_gatherbench.log
for f in gpurun_out/r7_*.log; do echo "== $f"; grep -E "^E:|staging|passed|failed" $f | sed -n '1p;2p'; done
