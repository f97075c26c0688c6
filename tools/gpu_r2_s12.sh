#!/bin/bash
mkdir -p gpurun_out
echo "== sorted tests"; timeout 1500 python -m pytest tests/test_gpu_bwd_sorted.py -x -q --timeout 900 2>&1 | tail -4
for c in 2 5; do timeout 600 python tools/bwd_ab2.py --cfg $c --reps 5 2>&1 | head -3; done
timeout 600 python tools/bwd_ab2.py --cfg 2 --batch 2 --dist uniform --reps 5 2>&1 | head -3
timeout 600 python tools/bwd_ab2.py --cfg 2 --batch 16 --dist uniform --reps 3 2>&1 | head -3
echo "== full gpu suite"; timeout 2400 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -4
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | cut -c1-330
