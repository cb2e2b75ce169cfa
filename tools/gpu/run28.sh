#!/bin/bash
# run 28: symmetrize parity + timing at box 256
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_reco_oracle.py -m gpu -q --tb=short -k "symmetrize or norm" > gpurun_out/r28_pytest.log 2>&1
grep -E "passed|failed|^E  " gpurun_out/r28_pytest.log | cut -c1-400
python - > gpurun_out/r28_symm_time.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np
from thunder_b200 import capi
from oracle import refapi
c = capi.Context(0)
c.reco_alloc(0, 512)
for g in ("C4", "D2", "T", "O"):
    R = refapi.symmetry_elements(g)
    c.symmetrize(0, R, 255.0)
    c.synchronize(); t = time.perf_counter()
    c.symmetrize(0, R, 255.0)
    c.synchronize(); print(g, len(R), "elements: %.1f ms at 512^3 (box 256, pf 2)" % ((time.perf_counter() - t) * 1e3))
c.close()
PY
cat gpurun_out/r28_symm_time.log
