set -x
python -m pytest tests/test_reco_oracle.py -m gpu -q --tb=short > gpurun_out/r18_pytest_reco.log 2>&1
grep -E "^E  |Error|assert|passed|failed" gpurun_out/r18_pytest_reco.log | cut -c1-250 | head -20
timeout 1500 python bench.py > gpurun_out/r18_bench.log 2> gpurun_out/r18_bench.err
tail -2 gpurun_out/r18_bench.log | cut -c1-3000; tail -3 gpurun_out/r18_bench.err
