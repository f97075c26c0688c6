#!/bin/bash
# Second GPU session: RED/TMA-reduce microbenchmarks, full variant sweep, other configs, ncu captures.
mkdir -p gpurun_out
echo "== microbench v2"; timeout 300 ./bm2f_b200/msda_microbench > gpurun_out/microbench2.txt 2>&1; grep -E "scaling|tma|d2d" gpurun_out/microbench2.txt
echo "== bf16 + host tests"; timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -k "bf16 or host_entry or staging_variants" 2>&1 | tail -5
echo "== full sweep"; timeout 900 python tools/sweep.py > gpurun_out/sweep_full.txt 2>&1; tail -80 gpurun_out/sweep_full.txt
for c in 1 3 4 5; do echo "== bench cfg $c"; timeout 600 python bench.py --cfg $c --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err; cat gpurun_out/bench_cfg$c.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['config']['workload'], d['value'], 'img/s', d['ms_per_step'],'ms', 'roof',d['roofline']['frac'], 'gather',{k:round(v['line_traffic_GBs']) for k,v in d['gather'].items() if isinstance(v,dict) and 'line_traffic_GBs' in v}, 'cpu',d['cpu_baseline'] and d['cpu_baseline']['value'], 'refcuda', d['reference_cuda'] and d['reference_cuda']['value'])"; done
echo "== ncu launch list"
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-ref-cuda > gpurun_out/ncu_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-ref-cuda > gpurun_out/ncu_bench.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/launches_r01.csv
echo "== ncu full"
python tools/ncu_target.py 16 > gpurun_out/ncu_plain_target.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 2 -c 2 -o gpurun_out/prof_r01 -f python tools/ncu_target.py 16 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out/
