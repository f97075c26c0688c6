"""The reference's OWN CUDA op (oracle/_ref/msda_reference_cuda, compiled in place from /root/reference by
oracle/build_ref.py) against the oracle and against this repository's kernels on the same inputs, so that a
difference between the two GPU implementations can be attributed (baseline 3).

Reference entry points: ms_deform_attn_forward / ms_deform_attn_backward
(/root/reference/mask2former/modeling/pixel_decoder/ops/src/vision.cpp:18-21, src/cuda/ms_deform_attn_cuda.cu:25-158).
The measured errors are written to gpurun_out/r02_reference_op_parity.json when that directory exists.
"""
import json
import os

import numpy as np
import pytest
import torch

from bm2f_b200 import cabi
from bm2f_b200 import workloads as W
from oracle import build_ref
from tests.helpers import rel_err, smooth_mask
from tests.test_gpu_parity import SMALL_LEVELS, oracle_ref, run_cabi

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref_op():
    mod = build_ref.load()
    if mod is None:
        pytest.skip("oracle/_ref/msda_reference_cuda was never built (needs /root/reference at build time)")
    return mod


def run_reference_op(mod, inp, im2col_step=128):
    dev = torch.device("cuda:0")
    v, lo, at, go = (inp[k].to(dev).contiguous() for k in ("value", "loc", "attn", "grad_out"))
    sh, st = inp["shapes"].to(dev), inp["start"].to(dev)
    out = mod.ms_deform_attn_forward(v, sh, st, lo, at, im2col_step)
    gv, gl, ga = mod.ms_deform_attn_backward(v, sh, st, lo, at, go.view_as(out), im2col_step)
    torch.cuda.synchronize()
    return dict(out=out.double().cpu().numpy(), grad_value=gv.double().cpu().numpy(),
                grad_loc=gl.double().cpu().numpy(), grad_attn=ga.double().cpu().numpy())


CASES = [("small_ragged", lambda: W.make_inputs(SMALL_LEVELS, 3, seed=77)),
         ("cfg1_model", lambda: W.workload_inputs(1, batch=1)),
         ("cfg1_uniform", lambda: W.workload_inputs(1, batch=1, dist="uniform")),
         ("cfg2_b2", lambda: W.workload_inputs(2, batch=2)),
         ("cfg5_b4", lambda: W.workload_inputs(5, batch=4))]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_reference_op_vs_oracle_vs_ours(name, make, ref_op, built):
    inp = make()
    want = oracle_ref(inp)
    ref = run_reference_op(ref_op, inp)
    ours = run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"])
    ok = smooth_mask(inp["loc"].numpy(), inp["shapes"].numpy())
    report = {}
    for tag, got in (("reference_cuda", ref), ("ours", ours)):
        report[tag] = dict(
            fwd_max_abs=float(np.abs(got["out"] - want["out"]).max()),
            grad_value_rel=float(rel_err(got["grad_value"], want["grad_value"])),
            grad_attn_rel=float(rel_err(got["grad_attn"], want["grad_attn"])),
            grad_loc_rel=float(rel_err(got["grad_loc"] * ok, want["grad_loc"] * ok)))
    report["ours_vs_reference_cuda"] = dict(
        fwd_max_abs=float(np.abs(ours["out"] - ref["out"]).max()),
        grad_value_rel=float(rel_err(ours["grad_value"], ref["grad_value"])),
        grad_attn_rel=float(rel_err(ours["grad_attn"], ref["grad_attn"])),
        grad_loc_rel=float(rel_err(ours["grad_loc"] * ok, ref["grad_loc"] * ok)))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        path = os.path.join(out_dir, "r02_reference_op_parity.json")
        allr = json.load(open(path)) if os.path.exists(path) else {}
        allr[name] = report
        with open(path, "w") as f:
            json.dump(allr, f, indent=1, sort_keys=True)
    scale = max(1.0, float(np.abs(want["out"]).max()))
    # both GPU implementations meet the north-star bar against the oracle ...
    for tag in ("reference_cuda", "ours"):
        r = report[tag]
        assert r["fwd_max_abs"] <= 1e-5 * scale, (tag, r)
        assert max(r["grad_value_rel"], r["grad_attn_rel"], r["grad_loc_rel"]) <= 1e-4, (tag, r)
    # ... and therefore each other
    r = report["ours_vs_reference_cuda"]
    assert r["fwd_max_abs"] <= 2e-5 * scale and max(r["grad_value_rel"], r["grad_attn_rel"], r["grad_loc_rel"]) <= 2e-4, r
