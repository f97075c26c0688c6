"""Drop-in proof at the reference's own files: the reference's ops/test.py, ops/functions/ms_deform_attn_func.py and
ops/modules/ms_deform_attn.py run UNCHANGED (byte-identical copies staged by oracle/build_ref.py:stage_ops_py into the
git-ignored oracle/_ref/ops_unmodified/, digests pinned in tests/golden/ref_ops_sha256.json) on top of this
repository's `MultiScaleDeformableAttention` extension.

  * /root/reference/mask2former/modeling/pixel_decoder/ops/test.py:84-89 — its three checks (fp64 forward, fp32 forward,
    fp64 gradcheck for D in 30, 32, 64, 71, 1025, 2048, 3096) must all print "True";
  * the reference's MSDeformAttn module and this repository's module, same parameters, same inputs: same outputs and
    gradients.
"""
import hashlib
import json
import os
import subprocess
import sys

import pytest
import torch

from oracle import build_ref

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "bm2f_b200")


@pytest.fixture(scope="module")
def staged(built):
    d = build_ref.stage_ops_py()
    if d is None or not os.path.isdir(os.path.join(d, "ops")):
        pytest.skip("oracle/_ref/ops_unmodified was never staged (needs /root/reference at build time)")
    pinned = json.load(open(build_ref.MANIFEST))
    for rel, digest in pinned.items():
        got = hashlib.sha256(open(os.path.join(d, "ops", rel), "rb").read()).hexdigest()
        assert got == digest, f"{rel} is not the reference's file"
    return d


def test_reference_test_py_runs_unchanged(staged):
    """python test.py, exactly as the reference's README says, with our extension as MultiScaleDeformableAttention"""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([PKG, env.get("PYTHONPATH", "")])
    r = subprocess.run([sys.executable, "test.py"], cwd=os.path.join(staged, "ops"), env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("* ")]
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "r02_reference_test_py.txt"), "w") as f:
            f.write(r.stdout)
    assert r.returncode == 0, r.stdout[-3000:]
    assert len(lines) == 9, r.stdout[-3000:]                 # 2 forward checks + 7 gradchecks
    assert all(ln.startswith("* True") for ln in lines), "\n".join(lines)


def test_reference_module_on_our_extension_equals_our_module(staged):
    """the reference's MSDeformAttn (its Python, our native op) vs bm2f_b200's MSDeformAttn with the same parameters"""
    import bm2f_b200
    MSDA = bm2f_b200.load_extension()          # registers this repository's build as `MultiScaleDeformableAttention`
    assert os.path.dirname(os.path.abspath(MSDA.__file__)) == PKG          # ours, not the reference build
    # `ops` must resolve to the staged reference package (a namespace package, no __init__.py, like in the reference
    # tree) and not to bm2f_b200/ops: only the staged directory goes on sys.path
    assert "ops" not in sys.modules
    sys.path.insert(0, staged)
    try:
        from ops.modules import MSDeformAttn as RefModule                   # staged, unmodified
    finally:
        sys.path.remove(staged)
    assert os.path.abspath(sys.modules[RefModule.__module__].__file__).startswith(os.path.abspath(staged))
    from bm2f_b200 import workloads as W
    from bm2f_b200.ops.modules import MSDeformAttn

    from bm2f_b200 import cabi
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    # Kink-free by construction, so that two fp32 implementations cannot land on different sides of a bilinear kink
    # (where the one-sided derivatives legitimately differ): level ratios are powers of two, the offset bias has
    # fractional part 0.3 (pixel coordinates then have fractional parts .05 / .175 / .3 / .425 / .55 / .675 / .8 / .925)
    # and the learned part of the offsets stays below 0.04 px.
    levels = ((8, 12), (16, 24), (32, 48))
    ours = MSDeformAttn(256, 3, 8, 4).to(dev)
    theirs = RefModule(256, 3, 8, 4).to(dev)
    with torch.no_grad():
        for name, prm in ours.named_parameters():
            if name == "sampling_offsets.weight":
                prm.copy_(4e-4 * torch.randn_like(prm))
            elif name == "sampling_offsets.bias":
                prm.add_(0.3)
            else:
                prm.add_(0.02 * torch.randn_like(prm))
    theirs.load_state_dict(ours.state_dict())                               # same keys: a zoo checkpoint loads either way
    shapes, start = W.level_tensors(levels, dev)
    S = int(shapes.prod(1).sum())
    ref_pts = W.reference_points(levels, 2).to(dev)
    src = torch.randn(2, S, 256, device=dev)
    pos = torch.randn(2, S, 256, device=dev)
    res = []
    for mod in (theirs, ours):
        s = src.clone().requires_grad_(True)
        n0 = cabi.lib().bm2f_msda_launch_count()
        out = mod(s + pos, ref_pts, s, shapes, start, None)
        out.square().sum().backward()
        torch.cuda.synchronize()
        # the reference module swallows native errors and falls back to torch (ms_deform_attn.py:116-121): make sure
        # the native kernels of THIS repository really ran underneath it
        assert cabi.lib().bm2f_msda_launch_count() - n0 >= 2
        res.append((out.detach(), s.grad, [prm.grad.clone() for prm in mod.parameters()]))
    (o_t, g_t, p_t), (o_o, g_o, p_o) = res
    scale = o_t.abs().max().item()
    assert (o_t - o_o).abs().max().item() <= 2e-5 * max(scale, 1.0)
    assert (g_t - g_o).abs().max().item() <= 1e-4 * g_t.abs().max().item()
    for (name, _), a, b in zip(ours.named_parameters(), p_t, p_o):
        assert (a - b).abs().max().item() <= 1e-4 * max(a.abs().max().item(), 1e-6), name
