#!/bin/bash
# two GPUs: the NCCL correctness test, the default bench through torch.distributed.run, config-1 shape and the 2D scan re-measured
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -p no:hypothesispytest > gpurun_out/r2_18_pytest_multi.log 2>&1
grep -E "passed|failed|skipped|^FAILED|^E  " gpurun_out/r2_18_pytest_multi.log | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_18_bench_2gpu.log 2> gpurun_out/r2_18_bench_2gpu.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_18_bench_2gpu.log") if l.startswith("{")][-1])
    print("N=2: value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],1), "frac", round(j["roofline"]["frac"],3), "allreduce", j.get("allreduce"), j.get("allreduce_selfcheck"), "clk", j["clocks"]["sm_mhz"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_18_bench_2gpu.err").read()[-2000:])
PY
timeout 600 python bench.py --box 128 --particles 1000 --batch 1000 --mlr 25 --phases 8 --cpu-sample 64 > gpurun_out/r2_18_bench_config1.log 2> gpurun_out/r2_18_bench_config1.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_18_bench_config1.log") if l.startswith("{")][-1])
    print("config1: value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],2), "frac", round(j["roofline"]["frac"],3), "cpu", j["cpu_baseline"]["value"])
except Exception as e:
    print("config1 bench failed", e); print(open("gpurun_out/r2_18_bench_config1.err").read()[-1500:])
PY
timeout 600 python tools/kbench2d.py --images 2048 2>&1 | grep "^scan" | tee gpurun_out/r2_18_kbench2d.log
timeout 900 python bench.py --mode 2d --steps 2 --warmup 2 > gpurun_out/r2_18_bench_2d.log 2> gpurun_out/r2_18_bench_2d.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_18_bench_2d.log") if l.startswith("{")][-1])
    print("2d: value", round(j["value"],1), "e2e", j["e2e"] and round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],1), "cpu", j.get("cpu_baseline"))
except Exception as e:
    print("2d bench failed", e); print(open("gpurun_out/r2_18_bench_2d.err").read()[-1500:])
PY
