#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests/test_gpu_bwd_sorted.py -x -q --timeout 600 -k "pixel_owner or config_shapes" 2>&1 | tail -5
for c in 2 4 5; do timeout 600 python tools/bwd_ab2.py --cfg $c --reps 5 2>&1 | tee -a gpurun_out/r02_bwd_ab3.txt; done
timeout 600 python tools/bwd_ab2.py --cfg 2 --batch 2 --dist uniform --reps 5 2>&1 | tee -a gpurun_out/r02_bwd_ab3.txt
