set -x
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r12_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --particles 20000 --batch 2500 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r12_bench_2gpu.log 2> gpurun_out/r12_bench_2gpu.err
python bench.py --gpus 1 --particles 10000 --batch 2500 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r12_bench_1gpu.log 2> gpurun_out/r12_bench_1gpu.err
tail -4 gpurun_out/r12_pytest_multi.log; tail -2 gpurun_out/r12_bench_2gpu.log | cut -c1-400; tail -3 gpurun_out/r12_bench_2gpu.err; tail -1 gpurun_out/r12_bench_1gpu.log | cut -c1-300
