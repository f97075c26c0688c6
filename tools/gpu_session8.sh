#!/bin/bash
mkdir -p gpurun_out
echo "== parity subset"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "golden or config_shapes or job_shapes or other_level or bf16 or fused" 2>&1 | tail -4
for c in 2 1 3 4 5; do echo "== bench cfg $c"; timeout 600 python bench.py --cfg $c --steps 10 --warmup 3 --no-e2e > gpurun_out/bench8_cfg$c.json 2> gpurun_out/bench8_cfg$c.err; python -c "
import sys,json
d=json.loads(open('gpurun_out/bench8_cfg$c.json').read()); print(d['config']['workload'], round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms', 'hbm frac',round(d['roofline']['frac'],3), 'binding', round(d['roofline'].get('frac_of_binding_limit',0),3), 'cpu',d['cpu_baseline'] and round(d['cpu_baseline']['value'],2), 'refcuda', d['reference_cuda'] and round(d['reference_cuda']['value'],1))"; done
