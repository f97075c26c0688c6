#!/bin/bash
# First GPU session: facts, microbenchmarks, parity tests, variant sweep, bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
ls /root/reference > gpurun_out/ref_ls.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
echo "== microbench"; timeout 300 ./bm2f_b200/msda_microbench > gpurun_out/microbench.txt 2>&1; tail -40 gpurun_out/microbench.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 600 python tools/sweep.py --quick > gpurun_out/sweep_quick.txt 2>&1; tail -20 gpurun_out/sweep_quick.txt
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
