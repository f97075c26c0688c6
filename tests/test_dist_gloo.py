"""CPU, world_size 2 over gloo: the N>1 host logic of the hot path — batch sharding, the flat
projection-gradient bucket (DDP semantics) and max-over-ranks timing."""
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bm2f_b200.dist import GradBucket, max_over_ranks, projection_parameters, shard_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_batch_covers_every_image_once():
    for total in (0, 1, 2, 15, 16, 17, 32):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                first, count = shard_batch(total, world, r)
                seen += list(range(first, first + count))
            assert seen == list(range(total))
    assert shard_batch(16, 8, 3) == (6, 2)          # cfg 2 at 8 GPUs: 2 images per rank
    with pytest.raises(ValueError):
        shard_batch(4, 2, 2)


class _FakeAttn(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.sampling_offsets = torch.nn.Linear(256, 192)
        self.attention_weights = torch.nn.Linear(256, 96)
        self.value_proj = torch.nn.Linear(256, 256)
        self.output_proj = torch.nn.Linear(256, 256)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                   # same parameters on every rank
    layers = [_FakeAttn() for _ in range(6)]
    params = projection_parameters(layers)
    for i, p in enumerate(params):                         # rank-dependent gradients
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    bucket = GradBucket(params)
    bucket.all_reduce()
    mean_scale = sum(r + 1 for r in range(world)) / world
    ok = all(torch.allclose(p.grad, torch.full_like(p, mean_scale * (i + 1))) for i, p in enumerate(params))
    slow = max_over_ranks(10.0 + rank, torch.device("cpu"))
    q.put((rank, bucket.numel, ok, slow))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_bucket_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, numel, ok, slow in res:
        assert numel == 1_233_600                          # 4.93 MB fp32
        assert ok
        assert slow == 11.0                                # max over ranks


def test_reference_arm_under_torchrun_only_rank0_works():
    # bench.py --impl reference: rank 0 alone runs and prints, the other ranks exit 0 without work
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference",
           "--gpus", "2", "--steps", "1", "--warmup", "0", "--cfg", "5"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    import json
    d = json.loads(lines[0])
    staged = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ops_unmodified", "ops", "functions", "ms_deform_attn_func.py"))
    assert d["impl"] == "reference" and d["value"] > 0
    # the reference's own function when its byte-identical copy travelled with the tree, the restatement otherwise
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["n_gpus"] == 2
    # a real, bounded step: value = sample images / measured step time, and the step list fits the run
    n = d["sample_images_per_step"]
    assert abs(d["value"] - n / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    # `config` is built by the function the GPU arm uses: same keys, same values for the same command line
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench_mod)
    from bm2f_b200 import workloads as W
    import types
    graph = bench_mod.graph_enabled(types.SimpleNamespace(graph=None), W.WORKLOADS[5], 16)     # launch-bound rule of both arms
    want = bench_mod.workload_config(W.WORKLOADS[5], 2, 16, "strong", "default", graph)
    assert json.loads(json.dumps(want)) == d["config"]
    assert d["config"]["parallelism"] == "dp2" and d["config"]["batch_per_gpu"] == 16 and d["scaling"] == "strong"
