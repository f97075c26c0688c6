#!/bin/bash
# round 2, session 6: balanced pixel-owner backward (variants 10-12): parity, phases
mkdir -p gpurun_out
echo "== owner tests"; timeout 1200 python -m pytest tests/test_gpu_bwd_sorted.py -x -q --timeout 600 -k "pixel_owner or config_shapes" 2>&1 | tail -8
echo "== phases"; timeout 600 python tools/bwd_phases.py --cfg 2 --variants 8,10,11,12 2>&1 | tee gpurun_out/r02_bwd_phases_owner.txt
for m in 4 5 8; do timeout 300 python tools/bwd_phases.py --cfg 2 --variants 10 --margin $m 2>&1 | tail -1; done
