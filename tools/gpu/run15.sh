set -x
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r15_pytest.log
python tools/kbench.py 1024 256 1e-3 > gpurun_out/r15_kbench_1e-3.log 2>&1
python tools/kbench.py 1024 256 1.5e-5 > gpurun_out/r15_kbench_1.5e-5.log 2>&1
THB_INSERT_IMPL=2 python tools/kbench.py 1024 256 1e-3 > gpurun_out/r15_kbench_ins2_1e-3.log 2>&1
grep -E "^E   |^E  |passed|failed|Error" gpurun_out/r15_pytest.log | head -30
grep -E "^M:" gpurun_out/r15_kbench_*.log
