#!/bin/bash
mkdir -p gpurun_out
echo "== linear + encoder tests"; timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_encoder.py -m gpu -q --timeout 300 -x > gpurun_out/pytest_enc.log 2>&1; tail -25 gpurun_out/pytest_enc.log
echo "== encoder bench"; timeout 600 python tools/encoder_bench.py > gpurun_out/encoder_bench.txt 2>&1; tail -6 gpurun_out/encoder_bench.txt
