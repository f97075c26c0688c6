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
  roofline dominant kernel (backward) against the measured HBM copy peak; the gather-level view
           (L1/L2 line traffic) is reported next to it under "gather"
  cpu_baseline / --impl reference: the reference's CPU path (F.grid_sample formulation, restated in
           oracle/msda_oracle.py:torch_port_forward) on this box's host cores, bounded sample

Multi-GPU (torchrun, one rank per GPU): the batch is per-GPU (weak scaling, images are independent
— SURVEY.md §8e); the only collective is the NCCL all-reduce of the 4.93 MB projection-weight
gradient bucket of the training configs, overlapped on a side stream.
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


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_sample(wl, n_images, n_layers, threads):
    """Times the reference's CPU path (torch port) on `n_images` x `n_layers` image-layers and
    returns images/s for the full 6-layer stack, extrapolated linearly (layers are identical work)."""
    from oracle import msda_oracle as O
    torch.set_num_threads(threads)
    inp = W.make_inputs(wl.levels, n_images, seed=1234 + wl.cfg)
    bwd = wl.mode == "fwd+bwd"
    dt = torch.float32

    def once():
        if bwd:
            O.torch_port_forward_backward(inp["value"].to(dt), inp["shapes"], inp["loc"].to(dt), inp["attn"].to(dt),
                                          inp["grad_out"].to(dt))
        else:
            with torch.no_grad():
                O.torch_port_forward(inp["value"].to(dt), inp["shapes"], inp["loc"].to(dt), inp["attn"].to(dt))
    once()                                     # warm-up
    times = []
    for _ in range(n_layers):
        t0 = time.perf_counter(); once(); times.append(time.perf_counter() - t0)
    per_layer = statistics.median(times)
    return n_images / (per_layer * wl.n_layers), per_layer


# --------------------------------------------------------------------------------------------
# reference arm: CPU only
# --------------------------------------------------------------------------------------------
def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n_img = 2 if wl.S < 8000 else 1
    t0 = time.perf_counter()
    step_times = []
    per_layer_all = []
    total = args.warmup + args.steps
    for i in range(total):
        ips, per_layer = cpu_reference_sample(wl, n_img, 1, threads)
        if i >= args.warmup:
            step_times.append(per_layer * wl.n_layers * wl.batch / n_img)    # full-batch 6-layer step, extrapolated
            per_layer_all.append(per_layer)
        if time.perf_counter() - t0 > 150:                                   # stay within a few minutes
            break
    per_layer = statistics.median(per_layer_all)
    ips = n_img / (per_layer * wl.n_layers)
    sample = (f"{n_img} image(s) x 1 layer {wl.mode} per step, {len(per_layer_all)} timed steps; images/s = "
              f"images / (median layer time x {wl.n_layers} layers)")
    line = {
        "impl": "reference", "metric": "MSDeformAttn %s images/s" % wl.mode, "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": len(per_layer_all), "warmup": args.warmup,
        "ms_per_step": 1e3 * statistics.median(step_times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "cfg": wl.cfg, "levels": wl.levels, "batch": wl.batch, "layers": wl.n_layers,
                   "mode": wl.mode, "note": "reference CPU path = ms_deform_attn_core_pytorch (F.grid_sample), "
                                            "restated in oracle/msda_oracle.py; the reference has no native CPU kernel"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample,
                         "cpu": cpu_model()},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


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

    import bm2f_b200
    from bm2f_b200 import build as b200_build
    from bm2f_b200 import cabi
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

    bwd = wl.mode == "fwd+bwd"
    vdt = torch.bfloat16 if wl.dtype == "bf16" else torch.float32
    nb = wl.batch                                   # per-GPU batch (weak scaling)
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

    grad_bucket = torch.zeros(PROJ_GRAD_ELEMS, device=dev) if (dist and bwd) else None
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
    if args.graph:
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
        for _ in range(3):                          # per-kernel times for the roofline: separate, un-graphed pass
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
    images_total = nb * world
    value = images_total / (ms_per_step * 1e-3)

    kms = {k: statistics.mean(a.elapsed_time(b) for a, b in v) for k, v in kernel_events.items() if v}

    # ---- end to end through the C ABI host entry (every rank drives its own GPU) -------------------
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(cabi, wl, base, nb, bwd, min(args.steps, 3), dist, world)
        except Exception as e:  # pragma: no cover
            log("e2e leg failed:", repr(e))
            e2e = {"value": None, "unit": "images/s", "error": repr(e)}

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
    roofline = {"bound": "hbm", "kernel": "msda_%s_fast_kernel (+grad_value zero-fill)" % dom if bwd else "msda_fwd_fast_kernel",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": facts.get("dram_bytes_per_launch", {}).get(f"cfg{wl.cfg}_{dom}"),
                "peak_source": peak_src, "ms_per_launch": kms[dom],
                "algorithmic_bytes_per_launch": alg_bytes}
    gather = {}
    for k in kms:
        gb = W.gather_bytes(wl.S, k, wl.dtype) * nb
        gather[k] = {"ms_per_launch": kms[k], "line_traffic_GBs": gb / (kms[k] * 1e-3) / 1e9,
                     "hbm_GBs": W.hbm_bytes(wl.S, k, wl.dtype) * nb / (kms[k] * 1e-3) / 1e9}
    peaks_g = facts.get("gather_peaks") or {}
    gather["measured_peaks"] = peaks_g or None
    if peaks_g:
        # forward: gathered 128-B lines vs the measured L1 line rate of the same access shape (8 lanes x 16 B);
        # backward: RED lines vs the measured red.global.add.v4.f32 rate (L2-side limit)
        if "fwd" in gather:
            gather["fwd"]["frac_of_measured_l1_gather_peak"] = gather["fwd"]["line_traffic_GBs"] / peaks_g["l1_resident_8x16B"]
        if "bwd" in gather:
            red_gbs = W.gather_bytes(wl.S, "fwd", wl.dtype) * nb / (kms["bwd"] * 1e-3) / 1e9
            gather["bwd"]["red_line_traffic_GBs"] = red_gbs
            gather["bwd"]["frac_of_measured_red_peak"] = red_gbs / peaks_g["red_v4_l2_resident"]
        # the HBM roofline is not the binding one for this op (gather/scatter demand is 15x the compulsory HBM
        # bytes): name the limit that binds the dominant kernel and the fraction reached
        if dom == "bwd" and "bwd" in gather:
            roofline["binding_limit"] = ("L2 red.global.add.v4.f32 line rate, measured %.0f GB/s "
                                         "(profiles/r01_microbench.txt)" % peaks_g["red_v4_l2_resident"])
            roofline["frac_of_binding_limit"] = gather["bwd"]["frac_of_measured_red_peak"]
        elif "fwd" in gather:
            roofline["binding_limit"] = ("L1 gather line rate of the 8 lanes x 16 B access shape, measured %.0f GB/s "
                                         "(profiles/r01_microbench.txt)" % peaks_g["l1_resident_8x16B"])
            roofline["frac_of_binding_limit"] = gather["fwd"]["frac_of_measured_l1_gather_peak"]

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

    # ---- CPU baseline (rank 0, N=1 only) --------------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        n_img = 2 if wl.S < 8000 else 1
        reps = 5
        ips, per_layer = cpu_reference_sample(wl, n_img, reps, threads)
        cpu = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "cpu": cpu_model(),
               "sample": f"{n_img} image(s) x 1 layer {wl.mode}, median of {reps} (1 warm-up); layer time "
                         f"{per_layer * 1e3:.0f} ms x {wl.n_layers} layers; torch {torch.__version__} "
                         f"{torch.get_num_threads()} threads"}

    line = {
        "metric": "MSDeformAttn %s images/s" % wl.mode, "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": {"workload": wl.name, "cfg": wl.cfg, "levels": wl.levels, "batch_per_gpu": nb,
                   "global_batch": images_total, "layers": wl.n_layers, "heads": 8, "head_dim": 32, "points": 4,
                   "mode": wl.mode, "parallelism": f"dp{world}",
                   "l2_policy": f"inputs larger than L2: {in_bytes / 1e6:.0f} MB of inputs per layer, 6 distinct "
                                "layer input sets, no flush needed",
                   "collective": "all-reduce of 4.93 MB projection-gradient bucket per step" if grad_bucket is not None else "none",
                   "tuning": args.tuning or "default", "cuda_graph": graph is not None},
        "roofline": roofline, "gather": gather, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
        "clocks": clocks, "reference_cuda": ref_cuda,
    }
    print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_e2e(cabi, wl, base, nb, bwd, steps, dist, world):
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
    return {"value": nb * world / (dt / steps), "unit": "images/s", "h2d_bytes_per_step": h2d,
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
    ap.add_argument("--graph", action="store_true",
                    help="replay one step (all layers, fwd+bwd) from a CUDA graph in the timed loop: for launch-bound "
                         "configs (cfg 1: one 512^2 image, 35 us of launch overhead per layer)")
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
