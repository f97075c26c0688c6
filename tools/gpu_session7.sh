#!/bin/bash
mkdir -p gpurun_out
echo "== linear tests"; timeout 300 python -m pytest tests/test_gpu_linear.py -m gpu -q --timeout 120 2>&1 | tail -6
echo "== linear bench"; timeout 300 python tools/linear_bench.py > gpurun_out/linear_bench2.txt 2>&1; cat gpurun_out/linear_bench2.txt
echo "== module bench"; timeout 600 python tools/module_bench.py > gpurun_out/module_bench2.txt 2>&1; tail -12 gpurun_out/module_bench2.txt
