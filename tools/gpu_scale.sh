#!/bin/bash
# multi-GPU bench lines: bash tools/gpu_scale.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
P=$((29500 + N))
echo "== strong N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_s$N.err | tee -a gpurun_out/r02_scaling_strong.jsonl | cut -c1-400; tail -2 gpurun_out/bench_s$N.err
echo "== strong, no graph replay N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+20)) bench.py --gpus $N --steps 10 --warmup 3 --no-graph --no-e2e 2> gpurun_out/bench_g$N.err | tee -a gpurun_out/r02_scaling_strong_nograph.jsonl | cut -c1-400
echo "== weak N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+40)) bench.py --gpus $N --steps 10 --warmup 3 --scaling weak --no-e2e 2> gpurun_out/bench_w$N.err | tee -a gpurun_out/r02_scaling_weak.jsonl | cut -c1-400
