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
