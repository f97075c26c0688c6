#!/bin/bash
mkdir -p gpurun_out
echo "== linear tests"; timeout 300 python -m pytest tests/test_gpu_linear.py -m gpu -q -x --timeout 120 > gpurun_out/pytest_linear.log 2>&1; tail -15 gpurun_out/pytest_linear.log
echo "== fused tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "fused or module" > gpurun_out/pytest_fused.log 2>&1; tail -15 gpurun_out/pytest_fused.log
echo "== module bench"; timeout 600 python tools/module_bench.py > gpurun_out/module_bench.txt 2>&1; tail -12 gpurun_out/module_bench.txt
