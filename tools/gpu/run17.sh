set -x
python -m pytest tests/test_reco_oracle.py -m gpu -q --tb=short > gpurun_out/r17_pytest_reco.log 2>&1
python -m pytest tests -m gpu -q --tb=line -k "not reco_oracle" 2>&1 | tail -12 > gpurun_out/r17_pytest_rest.log
grep -E "^E  |Error|assert|passed|failed" gpurun_out/r17_pytest_reco.log | head -30; tail -6 gpurun_out/r17_pytest_rest.log
