#!/bin/bash
mkdir -p gpurun_out
echo "== bench N=1"; timeout 900 python bench.py --steps 10 --warmup 3 --ddp-module 2> gpurun_out/bench_n1.err | tee gpurun_out/r02_bench_n1.jsonl | cut -c1-3000; tail -3 gpurun_out/bench_n1.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tee gpurun_out/r02_bench_reference.jsonl | cut -c1-1500
