#!/bin/bash
mkdir -p gpurun_out
echo "== owner tests"; timeout 1200 python -m pytest tests/test_gpu_bwd_sorted.py -x -q --timeout 600 -k "pixel_owner or config_shapes" 2>&1 | tail -8
echo "== phases"; timeout 600 python tools/bwd_phases.py --cfg 2 --variants 10,11,12 2>&1 | tee gpurun_out/r02_bwd_phases_owner3.txt
for m in 4 5 8; do timeout 300 python tools/bwd_phases.py --cfg 2 --variants 10 --margin $m 2>&1 | tail -1; done
echo "== ncu owner (variant 10)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_owner -s 1 -c 1 -o gpurun_out/prof_r02_bwd_owner_v10 -f python tools/bwd_phases.py --ncu 10,0 > gpurun_out/ncu_owner.log 2>&1; tail -2 gpurun_out/ncu_owner.log
