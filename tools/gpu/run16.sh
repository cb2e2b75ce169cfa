set -x
python -m pytest tests/test_gpu_hotpath.py -m gpu -q --tb=short -k "box256 or box128 or box512" > gpurun_out/r16_pytest_box.log 2>&1
python -m pytest tests -m gpu -q --tb=line -k "not box" 2>&1 | tail -15 > gpurun_out/r16_pytest_rest.log
python tools/kbench.py 1024 256 1.5e-5 2>&1 | grep "^M:" | tail -1 > gpurun_out/r16_kbench_M.log
grep -E "^E  |Error|assert|passed|failed" gpurun_out/r16_pytest_box.log | head -40; tail -5 gpurun_out/r16_pytest_rest.log; cat gpurun_out/r16_kbench_M.log
