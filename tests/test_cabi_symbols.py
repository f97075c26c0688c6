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
