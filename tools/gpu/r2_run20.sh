#!/bin/bash
# BASELINE config 4 on 4 GPUs with the round's final kernels: 20k particles, box 512, 4000 orientation samples (125 x 32 phases)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --box 512 --particles 20000 --batch 500 --phases 32 --steps 3 --warmup 3 --cpu-sample 16 > gpurun_out/r2_20_bench_config4_4gpu.log 2> gpurun_out/r2_20_bench_config4_4gpu.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_20_bench_config4_4gpu.log") if l.startswith("{")][-1])
    print("config4 N=4 value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],1), "frac", round(j["roofline"]["frac"],3), "E ms", round(j["roofline"]["avg_launch_ms"],1), "cpu", j["cpu_baseline"], "allreduce", j.get("allreduce"), j.get("allreduce_selfcheck"), "clk", j["clocks"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_20_bench_config4_4gpu.err").read()[-2500:])
PY
