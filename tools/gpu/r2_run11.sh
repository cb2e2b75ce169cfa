#!/bin/bash
# 2 GPUs: NCCL all-reduce correctness test + bench at N = 2 through torch.distributed.run
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -p no:hypothesispytest > gpurun_out/r2_11_pytest_multi.log 2>&1
grep -E "passed|failed|skipped|^E  " gpurun_out/r2_11_pytest_multi.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_11_bench_2gpu.log 2> gpurun_out/r2_11_bench_2gpu.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_11_bench_2gpu.log") if l.startswith("{")][-1])
    print("N=2 value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],1), "allreduce", j.get("allreduce"), "clk", j["clocks"]["sm_mhz"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_11_bench_2gpu.err").read()[-1500:])
PY
