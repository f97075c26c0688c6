#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/scale_n2b.json 2> gpurun_out/scale_n2b.err; echo "rc=$?"; tail -3 gpurun_out/scale_n2b.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/scale_n2b.json') if l.startswith('{')][-1])
print(2, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'launches', d['gpu_launches'], d['config']['collective'], d['clocks'])
PY
