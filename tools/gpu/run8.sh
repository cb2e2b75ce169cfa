set -x
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r8_pytest.log
for k in 1e-6 1.5e-5 7.6e-5; do
  python tools/kbench.py 1024 256 $k > gpurun_out/r8_kbench_t8x8_$k.log 2>&1
done
THB_TILE_W=8 THB_TILE_H=4 python tools/kbench.py 1024 256 1.5e-5 > gpurun_out/r8_kbench_t8x4_1.5e-5.log 2>&1
THB_TILE_W=16 THB_TILE_H=8 python tools/kbench.py 1024 256 1.5e-5 > gpurun_out/r8_kbench_t16x8_1.5e-5.log 2>&1
for f in gpurun_out/r8_*.log; do echo "== $f"; grep -E "^E:|staging|passed|failed" $f | sed -n '1p;2p'; done
