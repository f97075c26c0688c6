#!/bin/bash
# Round-end validation: smoke, full GPU suite, headline bench, all configs, ncu captures of the final kernels.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full gpu suite"; timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu_final.log 2>&1; tail -4 gpurun_out/pytest_gpu_final.log
echo "== bench default"; timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['roofline']['frac'], d['roofline'].get('frac_of_binding_limit'), d['e2e']['value'], d['reference_cuda']['value'], d['cpu_baseline']['value'])"
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-200
echo "== ncu launch list"
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-ref-cuda > gpurun_out/ncu_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-ref-cuda > gpurun_out/ncu_bench.log 2>&1
echo "rc=$?"
echo "== ncu full (sampling kernels, final defaults)"
python tools/ncu_target.py 16 > gpurun_out/ncu_plain_target.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 2 -c 2 -o gpurun_out/prof_r01_final -f python tools/ncu_target.py 16 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
