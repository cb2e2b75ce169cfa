set -x
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/r19_pytest.log
tail -6 gpurun_out/r19_pytest.log
