#!/bin/bash
mkdir -p gpurun_out
echo "== linear tests"; timeout 300 python -m pytest tests/test_gpu_linear.py -m gpu -q --timeout 120 2>&1 | tail -4
echo "== linear bench"; timeout 300 python tools/linear_bench.py > gpurun_out/linear_bench.txt 2>&1; cat gpurun_out/linear_bench.txt
echo "== module bench"; timeout 600 python tools/module_bench.py > gpurun_out/module_bench.txt 2>&1; tail -12 gpurun_out/module_bench.txt
echo "== full gpu test suite"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu_all.log 2>&1; tail -6 gpurun_out/pytest_gpu_all.log
echo "== bench default"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; python -c "
import json; d=json.load(open('gpurun_out/bench6.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['gather']['fwd'], d['gather']['bwd'], d['e2e']['value'], d['reference_cuda']['value'], d['cpu_baseline']['value'])"
echo "== ncu linear"
python tools/linear_bench.py 344064 > gpurun_out/lb_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:linear_tf32x3 -s 3 -c 2 -o gpurun_out/prof_linear_r01 -f python tools/linear_bench.py 344064 > gpurun_out/ncu_linear.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_linear.log
