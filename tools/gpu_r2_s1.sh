#!/bin/bash
# round 2, session 1: sorted backward — parity tests, A/B timing; sanity of the split build
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
echo "== sorted backward tests"; timeout 1200 python -m pytest tests/test_gpu_bwd_sorted.py -x -q --timeout 600 2>&1 | tail -15
echo "== A/B cfg2"; timeout 600 python tools/bwd_ab.py --cfg 2 2>&1 | tee gpurun_out/r02_bwd_ab_cfg2.txt
echo "== A/B cfg4"; timeout 600 python tools/bwd_ab.py --cfg 4 --reps 3 2>&1 | tee gpurun_out/r02_bwd_ab_cfg4.txt
echo "== A/B cfg5"; timeout 600 python tools/bwd_ab.py --cfg 5 --reps 3 2>&1 | tee gpurun_out/r02_bwd_ab_cfg5.txt
echo "== A/B cfg2 uniform"; timeout 600 python tools/bwd_ab.py --cfg 2 --batch 2 --reps 3 --dist uniform 2>&1 | tee gpurun_out/r02_bwd_ab_cfg2_uniform.txt
echo "== parity suite (split build sanity)"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -x -q --timeout 600 2>&1 | tail -5
