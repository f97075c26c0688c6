#!/bin/bash
# 2-GPU session: NCCL path of bench.py (weak scaling) + reference arm under torchrun.  Tight timeouts.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
echo "== N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err; echo "rc=$?"; tail -4 gpurun_out/scale_n2.err
echo "== N=1"; timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-ref-cuda > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "rc=$?"
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.loads([l for l in open(f'gpurun_out/scale_n{n}.json') if l.startswith('{')][-1])
        print(n, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'launches', d['gpu_launches'], d['config']['collective'], d['clocks'])
    except Exception as e: print(n, 'ERR', e)
PY
