"""TEST INFRASTRUCTURE (not part of the product path): golden vectors of the reference's whole pixel decoder.

Runs the UNMODIFIED reference classes `MSDeformAttnPixelDecoder` / `MSDeformAttnTransformerEncoderOnly` / `MSDeformAttn` /
`PositionEmbeddingSine` (mask2former/modeling/pixel_decoder/msdeformattn.py:165-358, ops/modules/ms_deform_attn.py,
transformer_decoder/position_encoding.py) on the CPU in THIS container — the module's own `except:` branch
(ops/modules/ms_deform_attn.py:116-121) sends the core op to `ms_deform_attn_core_pytorch` — and stores inputs'
recipe + outputs + gradients as tests/golden/decoder/pixel_decoder_tiny.npz.

Two third-party packages the reference file imports are absent from this image (and from /root/reference):
  detectron2 (v0.6; Mask2Former's INSTALL.md builds against it): `configurable`, `Conv2d`, `ShapeSpec`, `get_norm`,
      `SEM_SEG_HEADS_REGISTRY`
  fvcore (0.1.5): `nn.weight_init.c2_xavier_fill`
Their few used entry points are restated as stubs below from their published behaviour (Conv2d = conv -> norm ->
activation; get_norm("GN") = GroupNorm(32, C); c2_xavier_fill = kaiming_uniform(a=1), zero bias; the registry and
`configurable` are decorators that do not change the class when it is constructed with explicit arguments).

Parameters are NOT stored: both sides fill the state dict with `fill_state_dict` below (numpy Generator keyed by the
sorted parameter names), which keeps the fixture small.  The GPU test imports this file for that function only.

Kinks.  The network is piecewise smooth: ReLU (FFN, FPN output conv) and bilinear sampling have one-sided derivatives
at pre-activation 0 / integer pixel coordinates, and two float32 implementations may legitimately land on different
sides when a value is within ~1e-5 of a kink — one flipped unit then perturbs every upstream gradient (rank-1).  With
~10^5 units some always are that close, so `condition()` moves a handful of BIASES (`linear1.bias`,
`sampling_offsets.bias`, `layer_1.norm.bias`; each shifted by < 0.5) until every ReLU input is >= 2e-4 and every
sampling coordinate >= 1e-3 px away from its kink in the REFERENCE's forward.  The reference code is untouched; the
adjusted biases are stored in the fixture (`cond::<name>`) and loaded by both sides after `fill_state_dict`.

Usage (here, where /root/reference exists):  python oracle/gen_golden_decoder.py
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from collections import namedtuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "decoder", "pixel_decoder_tiny.npz")
REF_ROOT = "/root/reference"

CASE = dict(
    batch=2,
    image=(64, 96),
    input_shape={"res2": (32, 4), "res3": (256, 8), "res4": (512, 16), "res5": (320, 32)},   # name: (channels, stride)
    kwargs=dict(transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
                transformer_enc_layers=2, conv_dim=256, mask_dim=32, norm="GN",
                transformer_in_features=["res3", "res4", "res5"], common_stride=4),
    seed=20240,
)


def fill_state_dict(module, seed):
    """Deterministic parameters independent of torch's RNG: weights ~ N(0, 1/fan_in), biases ~ 0.1 N(0,1), norm scales
    1 + 0.1 N(0,1), level_embed ~ N(0,1); sampling-offset biases keep the module's compass initialisation (so samples
    spread over several pixels) plus noise."""
    rng = np.random.default_rng(seed)
    sd = module.state_dict()
    new = {}
    for name in sorted(sd.keys()):
        t = sd[name]
        shape = tuple(t.shape)
        z = rng.standard_normal(shape).astype(np.float32)
        if name.endswith("level_embed"):
            v = z
        elif name.endswith("sampling_offsets.bias"):
            v = t.detach().cpu().numpy().astype(np.float32) + 0.25 * z
        elif ".norm" in name or name.split(".")[-2].startswith("norm") or (name.startswith("input_proj") and ".1." in name):
            v = (1.0 + 0.1 * z) if name.endswith("weight") else 0.1 * z
        elif name.endswith("bias"):
            v = 0.1 * z
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            v = z / np.sqrt(fan_in)
        new[name] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    module.load_state_dict(new)
    return module


def make_features(case, device="cpu"):
    rng = np.random.default_rng(case["seed"] + 1)
    H, W = case["image"]
    feats = {}
    for name in sorted(case["input_shape"]):
        c, s = case["input_shape"][name]
        feats[name] = torch.from_numpy(rng.standard_normal((case["batch"], c, H // s, W // s)).astype(np.float32)).to(device)
    return feats


def make_probes(case, outputs, device="cpu"):
    """Fixed random cotangents: loss = sum_i <probe_i, output_i> over (mask_features, out0, multi_scale[0..2])."""
    rng = np.random.default_rng(case["seed"] + 2)
    return [torch.from_numpy(rng.standard_normal(tuple(o.shape)).astype(np.float32)).to(device) for o in outputs]


GRAD_KEYS = ["transformer.level_embed", "input_proj.0.0.weight", "input_proj.0.0.bias", "input_proj.0.1.weight",
             "input_proj.2.1.bias", "input_proj.1.0.weight", "transformer.encoder.layers.0.self_attn.sampling_offsets.bias",
             "transformer.encoder.layers.0.self_attn.value_proj.weight",
             "transformer.encoder.layers.1.self_attn.attention_weights.weight",
             "transformer.encoder.layers.0.norm1.weight", "transformer.encoder.layers.1.linear1.bias", "transformer.encoder.layers.1.linear2.bias",
             "adapter_1.weight", "layer_1.norm.weight",
             "mask_features.weight"]


def _best_shift(vals, period, lo, hi, steps=4001):
    """Shift d in [lo, hi] maximising the distance of every vals + d from the kink set {0} (period None) or the
    integers (period 1).  vals: (rows,) float64.  Returns (d, margin)."""
    grid = np.linspace(lo, hi, steps)
    x = vals[None, :] + grid[:, None]
    dist = np.abs(x) if period is None else np.abs(x - np.round(x))
    m = dist.min(1)
    i = int(np.argmax(m - 1e-9 * np.abs(grid)))          # smallest |shift| among (near-)ties
    return float(grid[i]), float(m[i])


def condition(dec, feats, relu_margin=2e-4, loc_margin=1e-3):
    """Move biases so that no ReLU input / sampling coordinate of the reference's forward sits next to a kink.
    Layer by layer in data-flow order, re-running the reference forward after each adjustment."""
    layers = dec.transformer.encoder.layers
    report, changed = [], {}

    def run_with_hook(module, fn):
        box = {}
        def hook(mod, args, out):            # must return None: a returned value would replace the module's output
            box["v"] = fn(mod, args, out)
        h = module.register_forward_hook(hook)
        with torch.no_grad():
            dec.forward_features(feats)
        h.remove()
        return box["v"]

    for li, layer in enumerate(layers):
        attn = layer.self_attn
        # 1. sampling coordinates: pix = ref * (W_l, H_l) - 0.5 + offset  (ops/modules/ms_deform_attn.py:101-109)
        def grab(mod, args, out):
            query, ref, _, shapes = args[0], args[1], args[2], args[3]
            off = mod.sampling_offsets(query)
            return off.double().numpy(), ref.double().numpy(), shapes.numpy()
        off, ref, shapes = run_with_hook(attn, grab)
        N, Lq, _ = off.shape
        M, L, P = attn.n_heads, attn.n_levels, attn.n_points
        off = off.reshape(N, Lq, M, L, P, 2)
        wh = np.stack([shapes[:, 1], shapes[:, 0]], -1).astype(np.float64)            # (L, 2) = (W, H)
        pix = ref[:, :, None, :, None, :] * wh[None, None, None, :, None, :] - 0.5 + off
        bias = attn.sampling_offsets.bias.detach().double().numpy().reshape(M, L, P, 2).copy()
        worst = 1.0
        for m in range(M):
            for l in range(L):
                for pt in range(P):
                    for c in range(2):
                        d, mg = _best_shift(pix[:, :, m, l, pt, c].reshape(-1), 1, -0.5, 0.5)
                        bias[m, l, pt, c] += d
                        worst = min(worst, mg)
        assert worst >= loc_margin, worst
        with torch.no_grad():
            attn.sampling_offsets.bias.copy_(torch.from_numpy(bias.reshape(-1).astype(np.float32)))
        report.append((f"layer {li} sampling coordinates", worst))
        # 2. FFN pre-activations (msdeformattn.py:116)
        pre = run_with_hook(layer.linear1, lambda mod, args, out: out.double().numpy())
        pre = pre.reshape(-1, pre.shape[-1])
        b = layer.linear1.bias.detach().double().numpy().copy()
        worst = 1.0
        for u in range(pre.shape[1]):
            d, mg = _best_shift(pre[:, u], None, -0.3, 0.3)
            b[u] += d
            worst = min(worst, mg)
        assert worst >= relu_margin, worst
        with torch.no_grad():
            layer.linear1.bias.copy_(torch.from_numpy(b.astype(np.float32)))
        report.append((f"layer {li} FFN ReLU inputs", worst))
    # 3. FPN output conv: conv -> GroupNorm -> ReLU (msdeformattn.py:268-277); the norm's bias is added last
    for idx in range(dec.num_fpn_levels):
        conv = getattr(dec, f"layer_{idx + 1}")
        pre = run_with_hook(conv.norm, lambda mod, args, out: out.double().numpy())
        b = conv.norm.bias.detach().double().numpy().copy()
        worst = 1.0
        for ch in range(pre.shape[1]):
            d, mg = _best_shift(pre[:, ch].reshape(-1), None, -0.3, 0.3)
            b[ch] += d
            worst = min(worst, mg)
        assert worst >= relu_margin, worst
        with torch.no_grad():
            conv.norm.bias.copy_(torch.from_numpy(b.astype(np.float32)))
        report.append((f"layer_{idx + 1} ReLU inputs", worst))
    # verify on a final forward (float32 rounding of the stored biases included) and collect what changed
    final = []
    for li, layer in enumerate(layers):
        pre = run_with_hook(layer.linear1, lambda mod, args, out: out.double().numpy())
        final.append((f"layer {li} FFN", float(np.abs(pre).min())))
        changed[f"transformer.encoder.layers.{li}.linear1.bias"] = layer.linear1.bias.detach().clone()
        changed[f"transformer.encoder.layers.{li}.self_attn.sampling_offsets.bias"] = \
            layer.self_attn.sampling_offsets.bias.detach().clone()
    for idx in range(dec.num_fpn_levels):
        conv = getattr(dec, f"layer_{idx + 1}")
        pre = run_with_hook(conv.norm, lambda mod, args, out: out.double().numpy())
        final.append((f"layer_{idx + 1}", float(np.abs(pre).min())))
        changed[f"layer_{idx + 1}.norm.bias"] = conv.norm.bias.detach().clone()
    for name, mg in report + final:
        print(f"  kink margin  {name:34s} {mg:.3e}")
    assert min(m for _, m in final) >= 0.5 * relu_margin
    return changed


def load_conditioned(module, npz):
    """Apply the fixture's adjusted biases (`cond::<name>`) on top of `fill_state_dict`."""
    sd = module.state_dict()
    for k in npz.files:
        if k.startswith("cond::"):
            sd[k[6:]] = torch.from_numpy(np.asarray(npz[k], dtype=np.float32)).to(sd[k[6:]].device)
    module.load_state_dict(sd)
    return module


def install_stubs():
    ShapeSpec = namedtuple("ShapeSpec", ["channels", "height", "width", "stride"], defaults=(None, None, None, None))

    class Conv2d(torch.nn.Conv2d):
        def __init__(self, *args, **kwargs):
            norm = kwargs.pop("norm", None)
            activation = kwargs.pop("activation", None)
            super().__init__(*args, **kwargs)
            self.norm = norm
            self.activation = activation

        def forward(self, x):
            x = torch.nn.functional.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
            if self.norm is not None:
                x = self.norm(x)
            if self.activation is not None:
                x = self.activation(x)
            return x

    def get_norm(norm, out_channels):
        if norm is None or norm == "":
            return None
        assert norm == "GN", norm
        return torch.nn.GroupNorm(32, out_channels)

    def configurable(init_func=None, **_):
        return init_func

    class _Registry:
        def register(self, obj=None):
            return obj if obj is not None else (lambda o: o)

    def c2_xavier_fill(module):
        torch.nn.init.kaiming_uniform_(module.weight, a=1)
        if module.bias is not None:
            torch.nn.init.constant_(module.bias, 0)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("detectron2")
    mod("detectron2.config", configurable=configurable)
    mod("detectron2.layers", Conv2d=Conv2d, ShapeSpec=ShapeSpec, get_norm=get_norm)
    mod("detectron2.modeling", SEM_SEG_HEADS_REGISTRY=_Registry())
    mod("fvcore")
    mod("fvcore.nn")
    sys.modules["fvcore.nn"].weight_init = mod("fvcore.nn.weight_init", c2_xavier_fill=c2_xavier_fill)
    # the compiled op is absent on the CPU: an empty module makes the reference take its own pure-torch branch
    sys.modules["MultiScaleDeformableAttention"] = types.ModuleType("MultiScaleDeformableAttention")
    # package shells so the reference's relative imports resolve without running mask2former/__init__.py
    for pkg, path in (("mask2former", "mask2former"), ("mask2former.modeling", "mask2former/modeling"),
                      ("mask2former.modeling.transformer_decoder", "mask2former/modeling/transformer_decoder"),
                      ("mask2former.modeling.pixel_decoder", "mask2former/modeling/pixel_decoder")):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF_ROOT, path)]
        sys.modules[pkg] = m
    return ShapeSpec


def main():
    if not os.path.isdir(REF_ROOT):
        raise SystemExit("reference tree not present: " + REF_ROOT)
    ShapeSpec = install_stubs()
    ref = importlib.import_module("mask2former.modeling.pixel_decoder.msdeformattn")
    torch.manual_seed(0)
    case = CASE
    shapes = {k: ShapeSpec(channels=c, stride=s) for k, (c, s) in case["input_shape"].items()}
    dec = ref.MSDeformAttnPixelDecoder(shapes, **case["kwargs"])
    fill_state_dict(dec, case["seed"])
    dec.train()
    feats = make_features(case)
    conditioned = condition(dec, feats)
    for v in feats.values():
        v.requires_grad_(True)
    mask_features, out0, multi = dec.forward_features(feats)
    outputs = [mask_features, out0] + list(multi)
    probes = make_probes(case, outputs)
    loss = sum((p * o).sum() for p, o in zip(probes, outputs))
    loss.backward()
    params = dict(dec.named_parameters())
    blob = {"mask_features": mask_features, "out0": out0}
    for i, m in enumerate(multi):
        blob[f"multi_scale_{i}"] = m
    for k, v in feats.items():
        blob[f"grad_feature_{k}"] = v.grad
    for k in GRAD_KEYS:
        blob["grad_param::" + k] = params[k].grad
    for k, v in conditioned.items():
        blob["cond::" + k] = v
    # the position embedding on its own (reference PositionEmbeddingSine, lowest level and an odd-sized one)
    pe = ref.PositionEmbeddingSine(128, normalize=True)
    blob["pos_2x3"] = pe(torch.zeros(1, 1, 2, 3))[0]
    blob["pos_7x5"] = pe(torch.zeros(1, 1, 7, 5))[0]
    blob["state_dict_keys"] = np.array(sorted(dec.state_dict().keys()))
    np.savez_compressed(OUT, **{k: (v.detach().numpy().astype(np.float32) if torch.is_tensor(v) else v)
                                for k, v in blob.items()})
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", "loss", float(loss))


if __name__ == "__main__":
    main()
