#!/usr/bin/env python
"""MSDeformAttn fwd+bwd throughput on B200 (BASELINE.json metric) — one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--cfg 2] [--impl ours|reference]

A *step* is one pass of the hot path over one batch: the 6 encoder layers' MSDeformAttn core op
(forward, and backward where the config trains) for every image of the batch, each layer on its
own input tensors.  Default workload = BASELINE.json configs[1]: COCO panoptic R50 1024x1024 LSJ,
8 heads x 3 levels x 4 points, batch 16, fp32, fwd+bwd.

  value    images/s, whole job, inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e      same metric through the C ABI's host-buffer entry: pinned host inputs copied in and all
           results copied out inside the timed region
  roofline dominant kernel (backward) against the measured HBM copy peak, plus SURVEY §8d's two-level
           bound max(bytes_hbm / BW_hbm, bytes_gather / BW_L2, bytes_red / BW_red) with the gather / RED
           peaks measured by bm2f_b200/msda_microbench ON THIS GPU IN THIS RUN (before the timed region)
  cpu_baseline / --impl reference: the reference's CPU path on this box's host cores, bounded sample
           (1 image of the batch, all 6 layers per step): the reference's OWN ms_deform_attn_core_pytorch
           from the byte-identical staged copy oracle/_ref/ops_unmodified (kind "reference") when it
           travelled with the tree, else the restatement oracle/msda_oracle.py (kind "port")

Multi-GPU (torchrun, one rank per GPU): the partition SURVEY.md §8e / BASELINE config 2 name — the
global batch (16) is split over the ranks, 16/8/4/2 images per GPU ("strong"); --scaling weak keeps
16 images per GPU.  Images are independent; the only collective is the NCCL all-reduce of the
4.93 MB projection-weight gradient bucket of the training configs: real gradients of six MSDeformAttn
modules (bm2f_b200.dist.GradBucket), launched on a side stream and overlapped with the step.  The
module-level step with one bucket per layer overlapped with the backward of the earlier layers is
reported next to it under "ddp_module".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bm2f_b200 import workloads as W  # noqa: E402

PROJ_GRAD_ELEMS = 1_233_600      # 4 Linear x 6 layers
FALLBACK_HBM_GBS = 6650.0        # /opt/skills/guides/B200_PROFILING.md


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples in the upper half of the observed clock range
        load = [x for x in sm if x >= 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": FALLBACK_HBM_GBS}, "fallback (B200_PROFILING.md)"


def profile_facts():
    """Numbers taken from committed profiles (ncu DRAM traffic, measured gather peaks)."""
    p = os.path.join(ROOT, "profiles", "facts.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return {}


def measure_gather_peaks(gpu_index: int):
    """L1 / L2 gather and L2 RED line rates of THIS GPU, measured now by the stand-alone microbenchmark
    (bm2f_b200/csrc/msda_microbench.cu --quick: 8 lanes x 16 B gathers from a 44 MB / 96 KB footprint,
    red.global.add.v4.f32 to a 44 MB footprint, device copy).  None when the binary is missing."""
    exe = os.path.join(ROOT, "bm2f_b200", "msda_microbench")
    if not os.path.exists(exe):
        return None
    env = dict(os.environ)
    vis = env.get("CUDA_VISIBLE_DEVICES")
    env["CUDA_VISIBLE_DEVICES"] = vis.split(",")[gpu_index] if vis else str(gpu_index)
    try:
        r = subprocess.run([exe, "--quick"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, env=env)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
        d = json.loads(line)
        d["source"] = "this run (bm2f_b200/msda_microbench --quick, same GPU, before the timed region)"
        return d
    except Exception as e:  # pragma: no cover
        log("microbenchmark failed:", repr(e))
        return None


def reference_cpu_function():
    """(callable, kind).  The reference's own CPU path, ms_deform_attn_core_pytorch
    (/root/reference/mask2former/modeling/pixel_decoder/ops/functions/ms_deform_attn_func.py:52-72), imported from the
    byte-identical copy oracle/build_ref.py staged under the git-ignored oracle/_ref/ops_unmodified (it travels to the
    GPU box; /root/reference does not) -> kind "reference".  Without it: the restatement in oracle/ -> kind "port"."""
    import importlib
    import types
    stage = os.path.join(ROOT, "oracle", "_ref", "ops_unmodified")
    fn = os.path.join(stage, "ops", "functions", "ms_deform_attn_func.py")
    if os.path.exists(fn):
        try:
            # the file imports the native module at the top and only calls it for CUDA tensors: on the CPU path a
            # stand-in of that name is enough (and keeps every kernel of this repository out of the reference arm)
            if "MultiScaleDeformableAttention" not in sys.modules:
                sys.modules["MultiScaleDeformableAttention"] = types.ModuleType("MultiScaleDeformableAttention")
            if stage not in sys.path:
                sys.path.insert(0, stage)
            mod = importlib.import_module("ops.functions.ms_deform_attn_func")
            return mod.ms_deform_attn_core_pytorch, "reference"
        except Exception as e:  # pragma: no cover
            log("staged reference function not importable:", repr(e))
    from oracle import msda_oracle as O
    return (lambda v, sh, loc, aw: O.torch_port_forward(v, sh, loc, aw)), "port"


def workload_config(wl, world, nb, scaling, tuning="default", cuda_graph=False):
    """`config` of the JSON line, identical in both arms (the driver compares them key by key): the workload, how it is
    partitioned, and the L2 policy, which is a property of the workload (its per-layer inputs exceed L2).  What is
    specific to an arm (the CPU sample size, notes) goes into `cpu_baseline.sample` / top-level keys instead."""
    esz = 2 if wl.dtype == "bf16" else 4
    in_bytes = nb * (wl.S * 8 * 32 * esz * (2 if wl.mode == "fwd+bwd" else 1) + wl.S * 8 * wl.L * 4 * 3 * 4)
    has_collective = wl.mode == "fwd+bwd" and wl.dtype == "f32" and world > 1
    return {"workload": wl.name, "cfg": wl.cfg, "levels": wl.levels, "global_batch": nb * world if scaling == "weak" else wl.batch,
            "layers": wl.n_layers, "heads": 8, "head_dim": 32, "points": 4, "mode": wl.mode, "parallelism": f"dp{world}",
            "batch_per_gpu": nb,
            "l2_policy": (f"inputs larger than L2: {in_bytes / 1e6:.0f} MB of inputs per layer, {wl.n_layers} distinct "
                          "layer input sets, no flush needed") if in_bytes > 126e6 else
                         (f"{in_bytes / 1e6:.0f} MB of inputs per layer, {wl.n_layers} distinct layer input sets rotate "
                          "(smaller than L2: launch-bound configuration, no explicit flush)"),
            "collective": ("all-reduce of the 4.93 MB projection-gradient bucket per step (real gradients of six "
                           "MSDeformAttn modules), side stream") if has_collective else "none",
            "tuning": tuning, "cuda_graph": bool(cuda_graph)}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_step(wl, fn, n_images, threads):
    """One bounded step of the reference's CPU path: `n_images` images of the batch through all wl.n_layers layers
    (forward, and backward through autograd where the config trains).  Returns the wall time of the step."""
    torch.set_num_threads(threads)
    bwd = wl.mode == "fwd+bwd"
    inp = cpu_reference_step.cache.get((wl.cfg, n_images))
    if inp is None:
        inp = W.make_inputs(wl.levels, n_images, seed=1234 + wl.cfg)
        cpu_reference_step.cache[(wl.cfg, n_images)] = inp
    t0 = time.perf_counter()
    for _ in range(wl.n_layers):
        if bwd:
            v, loc, aw = (inp[k].clone().requires_grad_(True) for k in ("value", "loc", "attn"))
            out = fn(v, inp["shapes"], loc, aw)
            out.backward(inp["grad_out"])
        else:
            with torch.no_grad():
                fn(inp["value"], inp["shapes"], inp["loc"], inp["attn"])
    return time.perf_counter() - t0


cpu_reference_step.cache = {}


def cpu_sample_images(wl):
    return 2 if wl.S < 8000 else 1


# --------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path, host cores only
# --------------------------------------------------------------------------------------------
def graph_enabled(args, wl, nb):
    """CUDA-graph replay of the step: explicit --graph / --no-graph, else on for launch-bound per-GPU steps"""
    return bool(args.graph) if args.graph is not None else nb * wl.S < 100000


def b200_shard(batch, world):
    """images of rank 0 when `batch` is split over `world` ranks (bm2f_b200.dist.shard_batch without importing torch.distributed)"""
    return batch // world + (1 if batch % world else 0)


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    fn, kind = reference_cpu_function()
    n_img = cpu_sample_images(wl)
    t_start = time.perf_counter()
    times = []
    for i in range(args.warmup + args.steps):
        dt = cpu_reference_step(wl, fn, n_img, threads)
        if i >= args.warmup:
            times.append(dt)
        if time.perf_counter() - t_start > 240 and times:                    # stay within a few minutes
            break
    step_s = sum(times) / len(times)
    ips = n_img / step_s
    sample = (f"each step = {n_img} image(s) of the batch x {wl.n_layers} layers {wl.mode} ({kind}: "
              f"ms_deform_attn_core_pytorch, F.grid_sample), {threads} torch threads; images/s = {n_img} / step time")
    scaling = args.scaling or ("strong" if world > 1 else "weak")
    nb = b200_shard(wl.batch, world) if scaling == "strong" else wl.batch
    cfg = workload_config(wl, world, nb, scaling, args.tuning or "default", graph_enabled(args, wl, nb))
    line = {
        "impl": "reference", "metric": "MSDeformAttn %s images/s" % wl.mode, "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * step_s,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "sample_images_per_step": n_img,
        "reference_note": "the reference has no native CPU kernel: its CPU path is ms_deform_attn_core_pytorch "
                          f"(ops/functions/ms_deform_attn_func.py:52-72); each step is a bounded sample of {n_img} image(s)",
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": kind, "sample": sample,
                         "cpu": cpu_model()},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def pin_to_gpu_numa(local: int):
    """Run this rank (and therefore its pinned allocations, first-touch) on the cores local to its GPU.  Returns a
    short description for the JSON line; a no-op when sysfs does not say or all GPUs sit on one node."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = open(base + "/numa_node").read().strip()
        cpus = open(base + "/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        ids &= set(os.sched_getaffinity(0))
        if ids:
            os.sched_setaffinity(0, ids)
        return {"pci": bdf, "numa_node": node, "local_cpulist": cpus, "pinned_cpus": len(ids)}
    except Exception as e:  # pragma: no cover
        return {"error": repr(e)}


def pcie_ceiling(dist, dev, nbytes=1 << 29, reps=4):
    """Both directions at once, pinned memory, EVERY rank at the same time (the ranks share the host's memory and PCIe
    root complexes): GB/s each way per GPU = the ceiling of the e2e leg on this box at this N."""
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def once(n):
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        s1.synchronize(); s2.synchronize()
        return time.perf_counter() - t0
    once(1)
    dt = once(reps)
    if dist:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return reps * nbytes / dt / 1e9


def run_ddp_module(wl, nb, dev, dist, world, modules, buckets, steps):
    """The training-config step at module level: six MSDeformAttn modules (fused sampling + tcgen05 projections, so the
    projection-weight gradients exist), forward + backward per layer, and one gradient bucket PER LAYER all-reduced on
    a side stream as soon as that layer's backward has been launched — it overlaps the backward of the next layer."""
    from bm2f_b200 import workloads as Wl
    shapes, start = Wl.level_tensors(wl.levels, dev)
    g = torch.Generator(device="cpu").manual_seed(99)
    src = torch.randn(nb, wl.S, 256, generator=g).to(dev)
    pos = (torch.randn(nb, wl.S, 256, generator=g) * 0.1).to(dev)
    gout = torch.randn(nb, wl.S, 256, generator=g).to(dev)
    ref_pts = Wl.reference_points(wl.levels, nb).to(dev)
    comm = torch.cuda.Stream()
    q = (src + pos)

    def step():
        for m, bk in zip(modules, buckets):
            for prm in bk.params:
                prm.grad = None
            qq = q.detach().requires_grad_(True)
            x = src.detach().requires_grad_(True)
            y = m(qq, ref_pts, x, shapes, start, None)
            y.backward(gout)
            if dist:
                bk.pack()
                comm.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(comm):
                    dist.all_reduce(bk.flat)
                    bk.flat.mul_(1.0 / world)
        torch.cuda.current_stream().wait_stream(comm)
    step(); step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    if dist:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this benchmark has no CPU path (use --impl reference "
                         "for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod

    numa = pin_to_gpu_numa(local)
    import bm2f_b200
    from bm2f_b200 import build as b200_build
    from bm2f_b200 import cabi
    from bm2f_b200 import dist as b200_dist
    # the native libraries normally travel with the tree; build them if this is a bare checkout
    # (local rank 0 compiles, the other ranks wait for the files)
    if not (os.path.exists(b200_build.LIB) and os.path.exists(b200_build.EXT)):
        if local == 0:
            log("[bench] native libraries missing: building (nvcc sm_100a + g++) ...")
            b200_build.build_all()
        else:
            t_wait = time.time()
            while not (os.path.exists(b200_build.LIB) and os.path.exists(b200_build.EXT)):
                if time.time() - t_wait > 900:
                    raise SystemExit("bench.py: timed out waiting for rank 0 to build the native libraries")
                time.sleep(2)
            time.sleep(5)
    MSDA = bm2f_b200.load_extension()
    if args.tuning:
        kv = dict(x.split("=") for x in args.tuning.split(","))
        cabi.set_default_tuning(cabi.make_tuning(**{k: int(v) for k, v in kv.items()}))

    # gather / RED peaks of this GPU, measured now (rank 0; the other ranks are idle meanwhile)
    peaks_run = measure_gather_peaks(local) if rank == 0 else None
    if dist:
        dist.barrier()

    bwd = wl.mode == "fwd+bwd"
    vdt = torch.bfloat16 if wl.dtype == "bf16" else torch.float32
    scaling = args.scaling or ("strong" if world > 1 else "weak")
    if scaling == "strong":
        # rank r owns images [r * N / G, (r + 1) * N / G) of the global batch
        _, nb = b200_dist.shard_batch(wl.batch, world, rank)
        if nb == 0:
            raise SystemExit(f"bench.py: global batch {wl.batch} leaves rank {rank} of {world} without images")
    else:
        nb = wl.batch                               # per-GPU batch fixed
    images_total = wl.batch if scaling == "strong" else nb * world
    step_arg = 128

    # ---- inputs: one seeded batch from the CPU generator, one copy per layer (rolled along batch) ----
    t0 = time.perf_counter()
    base = W.make_inputs(wl.levels, nb, seed=1234 + wl.cfg + 1000 * rank)
    shapes, start = base["shapes"].to(dev), base["start"].to(dev)
    layers = []
    for l in range(wl.n_layers):
        sh = l % nb
        layers.append(dict(
            value=torch.roll(base["value"], sh, 0).to(dev, vdt).contiguous(),
            loc=torch.roll(base["loc"], sh, 0).to(dev).contiguous(),
            attn=torch.roll(base["attn"], sh, 0).to(dev).contiguous(),
            grad_out=torch.roll(base["grad_out"], sh, 0).to(dev, vdt).contiguous()))
    log(f"[rank {rank}] inputs ready in {time.perf_counter() - t0:.1f}s: {wl.name} batch {nb} S={wl.S} "
        f"{wl.dtype} {wl.mode}")
    in_bytes = sum(t.numel() * t.element_size() for t in layers[0].values())

    # training configs: the projection-weight gradients of six MSDeformAttn modules (real values: one small
    # forward + backward through the module path fills them), one flat DDP-style bucket
    modules, grad_bucket, layer_buckets = None, None, None
    if bwd and wl.dtype == "f32" and (dist or args.ddp_module):
        from bm2f_b200.ops.modules import MSDeformAttn
        torch.manual_seed(7 + rank)
        modules = [MSDeformAttn(256, wl.L, 8, 4).to(dev) for _ in range(wl.n_layers)]
        sm_levels = ((4, 4), (8, 8), (16, 16))[:wl.L] if wl.L <= 3 else tuple((2 << i, 2 << i) for i in range(wl.L))
        sh_s, st_s = W.level_tensors(sm_levels, dev)
        S_s = sum(h * w for h, w in sm_levels)
        for m in modules:
            with torch.no_grad():
                m.sampling_offsets.weight.normal_(0, 0.01)
                m.attention_weights.weight.normal_(0, 0.05)
            xs = torch.randn(1, S_s, 256, device=dev, requires_grad=True)
            m(xs, W.reference_points(sm_levels, 1).to(dev), xs, sh_s, st_s, None).sum().backward()
        bucket_obj = b200_dist.GradBucket(b200_dist.projection_parameters(modules))
        assert bucket_obj.numel == PROJ_GRAD_ELEMS or wl.L != 3
        bucket_obj.pack()
        grad_bucket = bucket_obj.flat if dist else None
        layer_buckets = [b200_dist.GradBucket(b200_dist.projection_parameters([m])) for m in modules]
    comm_stream = torch.cuda.Stream() if grad_bucket is not None else None

    ev = lambda: torch.cuda.Event(enable_timing=True)
    kernel_events = {"fwd": [], "bwd": []}

    def step(record=False, collective=True):
        outs = None
        for L in layers:
            if record:
                a, b = ev(), ev(); a.record()
            out = MSDA.ms_deform_attn_forward(L["value"], shapes, start, L["loc"], L["attn"], step_arg)
            if record:
                b.record(); kernel_events["fwd"].append((a, b))
            if bwd:
                if record:
                    a, b = ev(), ev(); a.record()
                outs = MSDA.ms_deform_attn_backward(L["value"], shapes, start, L["loc"], L["attn"], L["grad_out"],
                                                    step_arg)
                if record:
                    b.record(); kernel_events["bwd"].append((a, b))
            else:
                outs = out
        if grad_bucket is not None and collective:
            # DDP-style: projection-weight gradients of the 6 layers, one flat bucket, side stream
            comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm_stream):
                dist.all_reduce(grad_bucket)
                grad_bucket.mul_(1.0 / world)
        return outs

    def sync_all():
        if comm_stream is not None:
            torch.cuda.current_stream().wait_stream(comm_stream)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()

    graph, graph_launches = None, 0
    if graph_enabled(args, wl, nb):
        # every launch of a step goes through the C ABI on the current stream: tensor maps are encoded on the host and
        # passed by value, the zero-fill is a memset node, nothing synchronises -> the whole step is capturable
        l0 = cabi.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step(collective=False)
        graph_launches = cabi.launch_count() - l0
        graph.replay()
        sync_all()

    def timed_step():
        if graph is None:
            return step(record=True)
        graph.replay()
        if grad_bucket is not None:
            comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm_stream):
                dist.all_reduce(grad_bucket)
                grad_bucket.mul_(1.0 / world)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = cabi.launch_count()
    e0, e1 = ev(), ev()
    sync_all()
    e0.record()
    for _ in range(args.steps):
        timed_step()
    if comm_stream is not None:
        torch.cuda.current_stream().wait_stream(comm_stream)
    e1.record()
    sync_all()
    elapsed_ms = e0.elapsed_time(e1)
    launches = cabi.launch_count() - launches0
    if graph is not None:
        launches = graph_launches * args.steps      # replayed launches do not pass through the launch counter
        for _ in range(6):                          # per-kernel times for the roofline: separate, un-graphed passes
            step(record=True, collective=False)
        torch.cuda.synchronize()
    # keep the GPU busy a little longer so the sampler certainly has samples under load
    if sampler:
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            step(collective=False)          # rank-local only: the other ranks are not in this loop
        torch.cuda.synchronize()
        clocks = sampler.stop()
    if dist:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = images_total / (ms_per_step * 1e-3)

    # un-graphed passes of a launch-bound step: the GPU waits for the host between launches, so single event pairs can
    # include host gaps -> median there; mean over the timed loop otherwise
    agg = statistics.median if graph is not None else statistics.mean
    kms = {k: agg([a.elapsed_time(b) for a, b in v]) for k, v in kernel_events.items() if v}

    # ---- end to end through the C ABI host entry (every rank drives its own GPU) -------------------
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(cabi, wl, base, nb, images_total, bwd, min(args.steps, 3), dist, world)
        except Exception as e:  # pragma: no cover
            log("e2e leg failed:", repr(e))
            e2e = {"value": None, "unit": "images/s", "error": repr(e)}

    if e2e and e2e.get("value"):
        try:
            ceil_gbs = pcie_ceiling(dist, dev)
            e2e["pcie_ceiling_GBs_each_way_per_gpu"] = ceil_gbs
            e2e["frac_of_pcie_ceiling"] = e2e["pcie_GBs_each_way"] / ceil_gbs
            e2e["pcie_ceiling_source"] = f"this run: {world} rank(s) copying both ways at once, pinned memory"
            e2e["numa"] = numa
        except Exception as e:  # pragma: no cover
            log("pcie ceiling probe failed:", repr(e))

    ddp_module = None
    if modules is not None and (dist or args.ddp_module):
        try:
            ms_mod = run_ddp_module(wl, nb, dev, dist, world, modules, layer_buckets, max(2, min(args.steps, 5)))
            ddp_module = {"value": images_total / (ms_mod * 1e-3), "unit": "images/s", "ms_per_step": ms_mod,
                          "what": "six MSDeformAttn modules fwd+bwd (fused sampling kernels + tcgen05 projections), one "
                                  "gradient bucket per layer all-reduced on a side stream while the next layer runs"
                                  if dist else "six MSDeformAttn modules fwd+bwd (no collective at N = 1)"}
        except Exception as e:  # pragma: no cover
            log("ddp_module leg failed:", repr(e))
            ddp_module = {"error": repr(e)}

    if rank != 0:
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    peaks, peak_src = measured_peaks()
    facts = profile_facts()
    dom = "bwd" if bwd else "fwd"
    alg_bytes = W.hbm_bytes(wl.S, dom, wl.dtype) * nb          # per launch (SURVEY §8d per image-layer x images)
    achieved = alg_bytes / (kms[dom] * 1e-3) / 1e9
    traffic = facts.get("dram_bytes_per_launch", {}).get(f"cfg{wl.cfg}_{dom}") if nb == wl.batch else None
    sorted_bwd = (bwd and wl.dtype == "f32" and nb * wl.S >= 32768 and "bwd=" not in (args.tuning or ""))   # msda_api.cu:choose_bwd_sorted
    kname = ("msda_bwd_sorted_kernel" if sorted_bwd else "msda_bwd_fast_kernel") + " (+grad_value zero-fill)" if bwd else "msda_fwd_fast_kernel"
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic, "peak_source": peak_src, "ms_per_launch": kms[dom],
                "algorithmic_bytes_per_launch": alg_bytes}
    gather = {}
    for k in kms:
        gb = W.gather_bytes(wl.S, k, wl.dtype) * nb
        gather[k] = {"ms_per_launch": kms[k], "line_traffic_GBs": gb / (kms[k] * 1e-3) / 1e9,
                     "hbm_GBs": W.hbm_bytes(wl.S, k, wl.dtype) * nb / (kms[k] * 1e-3) / 1e9}
    gather["measured_peaks"] = peaks_run
    if peaks_run:
        # SURVEY §8d's two-level bound, every peak measured on this GPU in this run: the op cannot run faster than
        # its compulsory HBM bytes at the HBM rate, its gathered corner lines at the L2 gather rate of the same
        # access shape (8 lanes x 16 B), and — backward, as long as every corner is one L2 reduction — its RED lines
        # at the red.global.add.v4.f32 rate
        l2g, l1g, red = (peaks_run["gather_8x16B_l2_44MB"], peaks_run["gather_8x16B_l1_64KB"],
                         peaks_run["red_8xv4f32_l2_44MB"])
        corner_bytes = W.gather_bytes(wl.S, "fwd", wl.dtype) * nb         # 48 lines per (query, head), once
        terms = {"hbm": alg_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3, "l2_gather": corner_bytes / (l2g * 1e9) * 1e3}
        # what the SM itself can take in: one 128-byte line per clock through the L1 data pipe (measured: an L1-resident
        # gather and a shared-memory gather of this shape both reach it, so a window staged in shared memory is no faster)
        terms["l1_data_pipe"] = corner_bytes / (l1g * 1e9) * 1e3
        lb = max(terms.values())
        roofline["lower_bound_ms"] = lb
        roofline["lower_bound_terms_ms"] = terms
        roofline["lower_bound_binding"] = max(terms, key=terms.get)
        roofline["frac_of_lower_bound"] = lb / kms[dom]
        if dom == "bwd":
            # not a bound of the op, a limit of every kernel that sends each bilinear corner to L2 as its own
            # reduction (the per-corner kernel sits on it; the anchor-sorted kernel merges corners that share an anchor,
            # the pixel-owner kernel escapes it and is issue-bound instead: profiles/README.md)
            roofline["one_red_per_corner_limit_ms"] = corner_bytes / (red * 1e9) * 1e3
        roofline["lower_bound_peaks"] = {"hbm_GBs": peaks["hbm_gbs"], "hbm_source": peak_src, "l2_gather_8x16B_GBs": l2g,
                                         "l1_gather_8x16B_GBs": l1g, "l2_red_v4f32_GBs": red,
                                         "gather_red_source": peaks_run["source"]}
        if "fwd" in gather:
            gather["fwd"]["frac_of_l1_gather_peak_this_run"] = gather["fwd"]["line_traffic_GBs"] / l1g
            gather["fwd"]["frac_of_l2_gather_peak_this_run"] = gather["fwd"]["line_traffic_GBs"] / l2g
        if "bwd" in gather:
            red_gbs = corner_bytes / (kms["bwd"] * 1e-3) / 1e9
            gather["bwd"]["red_line_traffic_GBs"] = red_gbs
            gather["bwd"]["frac_of_red_peak_this_run"] = red_gbs / red

    # ---- reference CUDA op, same inputs, same timing (only when oracle/_ref was built) ----------------
    ref_cuda = None
    if not args.no_ref_cuda and wl.dtype == "f32" and world == 1:
        try:
            from oracle import build_ref
            REF = build_ref.load()
        except Exception as e:  # pragma: no cover
            REF = None
            log("reference CUDA op not loadable:", e)
        if REF is not None:
            def ref_step():
                for L in layers:
                    REF.ms_deform_attn_forward(L["value"], shapes, start, L["loc"], L["attn"], step_arg)
                    if bwd:
                        REF.ms_deform_attn_backward(L["value"], shapes, start, L["loc"], L["attn"], L["grad_out"],
                                                    step_arg)
            ref_step(); ref_step()
            torch.cuda.synchronize()
            n = max(2, min(args.steps, 5))
            a, b = ev(), ev()
            a.record()
            for _ in range(n):
                ref_step()
            b.record(); torch.cuda.synchronize()
            ref_ms = a.elapsed_time(b) / n
            ref_cuda = {"value": nb / (ref_ms * 1e-3), "unit": "images/s", "ms_per_step": ref_ms, "steps": n,
                        "what": "reference's own CUDA op (Deformable-DETR kernels) compiled for sm_100a from "
                                "/root/reference in place (oracle/build_ref.py), same inputs, 1 GPU",
                        "speedup_ours": ref_ms / ms_per_step if world == 1 else None}

    # ---- CPU baseline (rank 0, N=1 only): same bounded step as --impl reference ---------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        fn, kind = reference_cpu_function()
        n_img = cpu_sample_images(wl)
        cpu_reference_step(wl, fn, n_img, threads)                      # warm-up
        ts = [cpu_reference_step(wl, fn, n_img, threads) for _ in range(3)]
        step_s = statistics.median(ts)
        cpu = {"value": n_img / step_s, "unit": "images/s", "cores": threads, "kind": kind, "cpu": cpu_model(),
               "sample": f"{n_img} image(s) of the batch x {wl.n_layers} layers {wl.mode} per step, median of 3 steps "
                         f"(1 warm-up), step {step_s * 1e3:.0f} ms; torch {torch.__version__} {torch.get_num_threads()} threads"}

    cfg = workload_config(wl, world, nb, scaling, args.tuning or "default", graph is not None)
    assert (cfg["collective"] != "none") == (grad_bucket is not None)
    line = {
        "metric": "MSDeformAttn %s images/s" % wl.mode, "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": cfg,
        "roofline": roofline, "gather": gather, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
        "clocks": clocks, "reference_cuda": ref_cuda, "ddp_module": ddp_module,
    }
    print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_e2e(cabi, wl, base, nb, images_total, bwd, steps, dist, world):
    """Same step through bm2f_msda_forward_backward_host: pinned host tensors in, results out."""
    code = cabi.DTYPE_BF16 if wl.dtype == "bf16" else cabi.DTYPE_F32
    vdt = torch.bfloat16 if wl.dtype == "bf16" else torch.float32
    pin = lambda t: t.contiguous().pin_memory()
    hv, hl, ha = pin(base["value"].to(vdt)), pin(base["loc"]), pin(base["attn"])
    hg = pin(base["grad_out"].to(vdt)) if bwd else None
    hs, hst = base["shapes"].contiguous(), base["start"].contiguous()
    ho = torch.empty(nb, wl.S, 256, dtype=vdt).pin_memory()
    hgv = torch.empty_like(hv).pin_memory() if bwd else None
    hgl = torch.empty_like(hl).pin_memory() if bwd else None
    hga = torch.empty_like(ha).pin_memory() if bwd else None
    dims = (nb, wl.S, 8, 32, wl.L, wl.S, 4)
    p = lambda t: t.data_ptr() if t is not None else 0

    def host_step():
        for _ in range(wl.n_layers):
            cabi.forward_backward_host(p(hv), p(hs), p(hst), p(hl), p(ha), p(hg), p(ho), p(hgv), p(hgl), p(hga),
                                       dims, code)
    host_step()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        host_step()
    dt = time.perf_counter() - t0
    if dist:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    nbytes = lambda *ts: sum(t.numel() * t.element_size() for t in ts if t is not None)
    h2d = wl.n_layers * nbytes(hv, hl, ha, hg)
    d2h = wl.n_layers * nbytes(ho, hgv, hgl, hga)
    return {"value": images_total / (dt / steps), "unit": "images/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * dt / steps, "steps": steps,
            "path": "bm2f_msda_forward_backward_host (C ABI), pinned host buffers, chunked copies pipelined over three slots",
            "pcie_GBs_each_way": max(h2d, d2h) / (dt / steps) / 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cfg", type=int, default=2, help="BASELINE config 1..5 (default 2 = headline)")
    ap.add_argument("--batch", type=int, default=None, help="override images per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tuning", default="", help="e.g. vec=4,staging=1,strip_w=16,rows=32,ctas_per_sm=2")
    ap.add_argument("--graph", dest="graph", action="store_const", const=True, default=None,
                    help="replay one step (all layers, fwd+bwd) from a CUDA graph in the timed loop.  Default: on when "
                         "the per-GPU step is launch-bound (fewer than 100 000 (image, query) rows per GPU: cfg 1, or "
                         "2-4 images per GPU in a strong-scaling run), off otherwise")
    ap.add_argument("--no-graph", dest="graph", action="store_const", const=False)
    ap.add_argument("--scaling", default="", choices=["", "strong", "weak"],
                    help="N > 1: strong (default) splits the config's global batch over the ranks (SURVEY section 8e: "
                         "16/8/4/2 images per GPU); weak keeps the config's batch on every GPU")
    ap.add_argument("--ddp-module", action="store_true",
                    help="also time the module-level step (six MSDeformAttn modules, per-layer gradient buckets) at N = 1")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    args = ap.parse_args()
    wl = W.WORKLOADS[args.cfg]
    if args.batch:
        from dataclasses import replace
        wl = replace(wl, batch=args.batch)
    if args.impl == "reference":
        return run_reference_arm(args, wl)
    return run_ours(args, wl)


if __name__ == "__main__":
    sys.exit(main())
