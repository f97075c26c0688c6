#!/bin/bash
# round 2, session 5: ncu (source-level) of the pixel-owner backward, variant 8
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_pixel -s 1 -c 1 -o gpurun_out/prof_r02_bwd_pixel_v8 -f python tools/bwd_phases.py --ncu 8,0 > gpurun_out/ncu_pixel.log 2>&1; tail -2 gpurun_out/ncu_pixel.log
