#!/bin/bash
mkdir -p gpurun_out
for c in 2 4 5 1; do timeout 600 python tools/bwd_ab2.py --cfg $c --reps 5 2>&1 | tee -a gpurun_out/r02_bwd_ab2.txt; done
timeout 600 python tools/bwd_ab2.py --cfg 2 --batch 2 --dist uniform --reps 5 2>&1 | tee -a gpurun_out/r02_bwd_ab2.txt
