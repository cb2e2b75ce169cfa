set -x
python -m pytest tests/test_reco_oracle.py tests/test_capi_load.py -m gpu -q --tb=short > gpurun_out/r20_pytest_reco.log 2>&1
grep -E "^E  |Error|assert|passed|failed" gpurun_out/r20_pytest_reco.log | cut -c1-250 | head -20
python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -6 > gpurun_out/r20_pytest_all.log; cat gpurun_out/r20_pytest_all.log
