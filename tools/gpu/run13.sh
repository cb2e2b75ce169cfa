set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r13_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r13_smoke.log 2>&1
tail -30 gpurun_out/r13_pytest.log; cat gpurun_out/r13_smoke.log | tail -3
