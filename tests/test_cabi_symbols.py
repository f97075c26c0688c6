"""CPU: the C-ABI library loads and exports every symbol include/bm2f_msda.h declares; the
argument checks that need no GPU behave like the reference's host wrapper."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "bm2f_msda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(bm2f_(?:msda|linear|add|zero|transpose|groupnorm|sine|conv3x3|fpn)_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported(built):
    from bm2f_b200 import cabi
    declared = _declared()
    assert sorted(cabi.SYMBOLS) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", cabi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (bm2f_(?:msda|linear|add|zero|transpose|groupnorm|sine|conv3x3|fpn)_[a-z0-9_]+)", out))
    assert set(declared) <= exported
    L = cabi.lib()
    for s in declared:
        assert getattr(L, s) is not None


def test_abi_version_and_build_info(built):
    from bm2f_b200 import cabi
    assert cabi.lib().bm2f_msda_abi_version() == 1
    assert b"sm_100a" in cabi.lib().bm2f_msda_build_info()


def test_library_contains_only_sm100a(built):
    from bm2f_b200 import cabi
    out = subprocess.run(["cuobjdump", "--list-elf", cabi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_im2col_step_rule(built):
    # reference: ms_deform_attn_cuda.cu:53-57 — step = min(batch, im2col_step); batch % step == 0
    from bm2f_b200 import cabi
    L = cabi.lib()
    assert L.bm2f_msda_check_im2col_step(16, 128) == 0
    assert L.bm2f_msda_check_im2col_step(256, 128) == 0
    assert L.bm2f_msda_check_im2col_step(130, 128) == -4
    assert "must divide" in cabi.last_error()
    assert L.bm2f_msda_check_im2col_step(6, 4) == -4
    assert L.bm2f_msda_check_im2col_step(4, 0) == -4


def test_null_and_bad_arguments_fail_without_gpu(built):
    from bm2f_b200 import cabi
    dims = (1, 4, 8, 32, 1, 4, 4)
    with pytest.raises(cabi.MSDAError, match="null"):
        cabi.forward(0, 0, 0, 0, 0, 0, dims)
    with pytest.raises(cabi.MSDAError, match="non-positive"):
        cabi.forward(16, 16, 16, 16, 16, 16, (0, 4, 8, 32, 1, 4, 4))
    with pytest.raises(cabi.MSDAError, match="dtype"):
        cabi.forward(16, 16, 16, 16, 16, 16, dims, dtype=7)


def test_extension_surface_and_cpu_error(built):
    # the drop-in module name and its two functions (ops/src/vision.cpp:18-21); CPU tensors raise
    # RuntimeError("Not implemented on the CPU") like ops/src/ms_deform_attn.h:43
    import torch

    import bm2f_b200
    m = bm2f_b200.load_extension()
    import MultiScaleDeformableAttention as MSDA
    assert MSDA is m
    assert callable(m.ms_deform_attn_forward) and callable(m.ms_deform_attn_backward)
    v = torch.zeros(1, 4, 8, 32)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        m.ms_deform_attn_forward(v, torch.tensor([[2, 2]]), torch.tensor([0]), torch.zeros(1, 4, 8, 1, 4, 2),
                                 torch.zeros(1, 4, 8, 1, 4), 128)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        m.ms_deform_attn_backward(v, torch.tensor([[2, 2]]), torch.tensor([0]), torch.zeros(1, 4, 8, 1, 4, 2),
                                  torch.zeros(1, 4, 8, 1, 4), torch.zeros(1, 4, 256), 128)


def test_fpn_entry_points_check_arguments_without_gpu(built):
    """FPN-tail entry points (include/bm2f_msda.h, msdeformattn.py:341-358): sizes and argument checks that run before
    anything touches a device; the torch-facing wrappers refuse CPU tensors like the rest of the module."""
    import ctypes

    import torch

    import bm2f_b200
    from bm2f_b200 import cabi
    L = cabi.lib()
    L.bm2f_conv3x3_workspace_bytes.restype = ctypes.c_size_t
    assert L.bm2f_conv3x3_workspace_bytes(256, 256) == 2 * 256 * 256 * 9 * 4      # hi + lo halves of the GEMM weights
    assert L.bm2f_conv3x3_workspace_bytes(0, 256) == 0
    vp, ci = ctypes.c_void_p, ctypes.c_int
    L.bm2f_conv3x3_forward.argtypes = [vp] * 4 + [ci] * 6 + [vp]
    L.bm2f_conv3x3_forward.restype = ci
    assert L.bm2f_conv3x3_forward(None, None, None, None, 1, 4, 4, 256, 256, 1, None) == -1          # BM2F_ERR_INVALID
    assert "null" in cabi.last_error()
    assert L.bm2f_conv3x3_forward(16, 16, 16, 16, 1, 4, 4, 128, 256, 1, None) == -2                  # BM2F_ERR_UNSUPPORTED
    assert "256 -> 256" in cabi.last_error()
    assert L.bm2f_conv3x3_forward(16, 16, 16, 16, 0, 4, 4, 256, 256, 1, None) == -1
    assert L.bm2f_conv3x3_forward(16, 16, 16, 16, 1, 4, 4, 256, 256, 2, None) == -1
    assert "split" in cabi.last_error()
    L.bm2f_conv3x3_set_variant.argtypes = [ci]
    assert L.bm2f_conv3x3_set_variant(1) == 0
    m = bm2f_b200.load_extension()
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        m.conv3x3_tokens_forward(torch.zeros(1, 6, 6, 256), torch.zeros(256, 256, 3, 3), 1)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        m.groupnorm_tokens_stats(torch.zeros(1, 4, 4, 256), 1e-5)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        m.fpn_upsample_backward(torch.zeros(1, 4, 4, 256), 2, 2)


def test_fused_fpn_tail_is_selected_only_for_the_standard_configuration(built):
    """Host logic of ops/functions/fpn_func.supported: GroupNorm(32) convs without bias, 256 channels, ReLU after the 3x3."""
    import torch

    from bm2f_b200.ops.functions import fpn_func
    from bm2f_b200.pixel_decoder import MSDeformAttnPixelDecoder, ShapeSpec
    shapes = {k: ShapeSpec(channels=c, stride=4 * 2 ** i) for i, (k, c) in enumerate(
        {"res2": 256, "res3": 512, "res4": 1024, "res5": 2048}.items())}
    kw = dict(transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024, transformer_enc_layers=1,
              conv_dim=256, mask_dim=256, transformer_in_features=["res3", "res4", "res5"], common_stride=4)
    gn = MSDeformAttnPixelDecoder(shapes, norm="GN", **kw)
    plain = MSDeformAttnPixelDecoder(shapes, norm="", **kw)
    x = torch.zeros(1, 256, 8, 8)
    # CPU tensors are never eligible (no CPU path); the same modules with a CUDA tensor are checked in the GPU suite
    assert not fpn_func.supported([x], gn.lateral_convs, gn.output_convs, gn.mask_features)

    class FakeCuda:
        is_cuda, dtype = True, torch.float32
    assert fpn_func.supported([FakeCuda()], gn.lateral_convs, gn.output_convs, gn.mask_features)
    assert not fpn_func.supported([FakeCuda()], plain.lateral_convs, plain.output_convs, plain.mask_features)
    half = FakeCuda()
    half.dtype = torch.float16
    assert not fpn_func.supported([half], gn.lateral_convs, gn.output_convs, gn.mask_features)
