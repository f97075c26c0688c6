#!/bin/bash
# round 2, session 4 (after container re-creation): re-measure everything the lost sessions measured
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
echo "== new tests"; timeout 1500 python -m pytest tests/test_gpu_bwd_sorted.py tests/test_gpu_reference_op.py tests/test_gpu_reference_dropin.py -q --timeout 900 2>&1 | tail -15
echo "== phases / variants cfg2"; timeout 900 python tools/bwd_phases.py --cfg 2 --variants 0,1,2,5,6,7,8,9 2>&1 | tee gpurun_out/r02_bwd_phases_cfg2.txt
echo "== A/B cfg2"; timeout 600 python tools/bwd_ab.py --cfg 2 --reps 5 2>&1 | tee gpurun_out/r02_bwd_ab_cfg2.txt
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tee gpurun_out/r02_bench_s4.txt | tail -3
echo "== parity suite"; timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --timeout 600 2>&1 | tail -5
