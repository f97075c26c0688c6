#!/bin/bash
# Round-end validation on one B200 (under gpurun): smoke, full GPU suite, headline bench + reference arm, the other
# BASELINE configs, ncu launch list and full capture of the default kernels, decoder / FPN / module benches.
bash "$(dirname "$0")/gpu_run.sh" smoke suite bench benchref allcfg launches ncusorted decbench fpnbench modbench
