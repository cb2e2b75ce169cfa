#!/bin/bash
# N = 8: the default bench through torch.distributed.run with the round's final kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2_27_bench_8gpu.log 2> gpurun_out/r2_27_bench_8gpu.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_27_bench_8gpu.log") if l.startswith("{")][-1])
    print("N=8: value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],1), "frac", round(j["roofline"]["frac"],3), "allreduce", j.get("allreduce"), j.get("allreduce_selfcheck"), "clk", j["clocks"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_27_bench_8gpu.err").read()[-2500:])
PY
