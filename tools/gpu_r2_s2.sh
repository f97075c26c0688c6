#!/bin/bash
# round 2, session 2: sorted backward — phase profile, launch variants, ncu (source-level) of the default variant
mkdir -p gpurun_out
echo "== sorted backward tests (quick)"; timeout 900 python -m pytest tests/test_gpu_bwd_sorted.py -x -q --timeout 600 2>&1 | tail -4
echo "== phases / variants cfg2"; timeout 900 python tools/bwd_phases.py --cfg 2 2>&1 | tee gpurun_out/r02_bwd_phases_cfg2.txt
echo "== ncu sorted (variant 0, lanes 4)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_sorted -s 1 -c 1 -o gpurun_out/prof_r02_bwd_sorted_v0 -f python tools/bwd_phases.py --ncu 0,4 > gpurun_out/ncu_sorted.log 2>&1; tail -2 gpurun_out/ncu_sorted.log
echo "== ncu sorted (variant 1, lanes 8)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_sorted -s 1 -c 1 -o gpurun_out/prof_r02_bwd_sorted_v1 -f python tools/bwd_phases.py --ncu 1,8 > gpurun_out/ncu_sorted1.log 2>&1; tail -2 gpurun_out/ncu_sorted1.log
echo "== new tests: reference op parity, unmodified drop-in"
timeout 1500 python -m pytest tests/test_gpu_reference_op.py tests/test_gpu_reference_dropin.py -q --timeout 900 2>&1 | tail -15
