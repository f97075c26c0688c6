#!/bin/bash
# round 2, session 3: pixel-owner backward — parity, phase profile, ncu
mkdir -p gpurun_out
echo "== sorted / pixel backward tests"; timeout 1500 python -m pytest tests/test_gpu_bwd_sorted.py -x -q --timeout 600 2>&1 | tail -6
echo "== phases / variants cfg2"; timeout 900 python tools/bwd_phases.py --cfg 2 2>&1 | tee gpurun_out/r02_bwd_phases2_cfg2.txt
echo "== margins, variant 8"; for m in 4 8; do timeout 300 python tools/bwd_phases.py --cfg 2 --variants 8 --margin $m 2>&1 | tail -1; done
echo "== cfg4 / cfg5, variant 8"; timeout 600 python tools/bwd_phases.py --cfg 4 --variants 0,8 --reps 3 2>&1 | tail -3; timeout 600 python tools/bwd_phases.py --cfg 5 --variants 0,8 --reps 3 2>&1 | tail -3
echo "== ncu pixel (variant 8)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_pixel -s 1 -c 1 -o gpurun_out/prof_r02_bwd_pixel_v8 -f python tools/bwd_phases.py --ncu 8,0 > gpurun_out/ncu_pixel.log 2>&1; tail -2 gpurun_out/ncu_pixel.log
echo "== drop-in module test"; timeout 600 python -m pytest tests/test_gpu_reference_dropin.py -q --timeout 600 2>&1 | tail -5
