set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"; free -g | head -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
python tools/kbench.py 1024 256 7.6e-5 > gpurun_out/r1_kbench_k7e-5.log 2>&1
python tools/kbench.py 1024 256 1e-6 > gpurun_out/r1_kbench_k1e-6.log 2>&1
python tools/kbench.py 1024 256 1e-3 > gpurun_out/r1_kbench_k1e-3.log 2>&1
timeout 900 python bench.py --particles 20000 --batch 2000 --steps 2 --warmup 1 > gpurun_out/r1_bench_small.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:expect_local -s 1 -c 1 -o gpurun_out/r1_prof_E python tools/kbench.py 296 256 7.6e-5 > gpurun_out/r1_ncuE.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:insert_kernel -s 1 -c 1 -o gpurun_out/r1_prof_M python tools/kbench.py 296 256 7.6e-5 > gpurun_out/r1_ncuM.log 2>&1
tail -3 gpurun_out/*.log
