"""ctypes binding of include/bm2f_msda.h — the same C ABI the torch extension calls.

Used by the parity tests ("-m gpu tests call through the C-ABI"), by bench.py (tuning sweeps,
launch counter, host-buffer end-to-end entry) and as the template for non-Python callers
(INTEGRATION.md).  Pointers are raw integers (``tensor.data_ptr()``); nothing here touches torch.
"""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libbm2f_msda.so")

DTYPE_F32, DTYPE_F64, DTYPE_BF16 = 0, 1, 2

# every symbol include/bm2f_msda.h declares (tests/test_cabi_symbols.py checks the export list)
SYMBOLS = (
    "bm2f_msda_abi_version", "bm2f_msda_build_info", "bm2f_msda_last_error", "bm2f_msda_launch_count",
    "bm2f_msda_set_default_tuning", "bm2f_msda_debug_phase_profile", "bm2f_msda_check_im2col_step", "bm2f_msda_forward",
    "bm2f_msda_backward", "bm2f_msda_forward_backward_host", "bm2f_msda_release_host_workspace", "bm2f_msda_fused_supported",
    "bm2f_msda_fused_forward", "bm2f_msda_fused_backward", "bm2f_msda_fused_forward_packed",
    "bm2f_msda_fused_backward_packed", "bm2f_linear_workspace_bytes", "bm2f_linear_set_tuning",
    "bm2f_linear_forward",
    "bm2f_linear_backward_input", "bm2f_linear_backward_weight", "bm2f_linear_relu_forward",
    "bm2f_linear_backward_input_masked", "bm2f_linear_backward_input_accumulate", "bm2f_add_layernorm_forward",
    "bm2f_add_layernorm_backward", "bm2f_zero_masked_rows", "bm2f_transpose_batched",
    "bm2f_groupnorm_tokens_workspace_bytes", "bm2f_groupnorm_tokens_forward", "bm2f_groupnorm_tokens_backward",
    "bm2f_sine_position_embedding", "bm2f_conv3x3_workspace_bytes", "bm2f_conv3x3_set_variant", "bm2f_conv3x3_forward", "bm2f_conv3x3_backward_input",
    "bm2f_conv3x3_backward_weight", "bm2f_groupnorm_tokens_stats", "bm2f_fpn_merge_forward", "bm2f_fpn_upsample_backward",
    "bm2f_groupnorm_relu_tokens_apply", "bm2f_groupnorm_relu_tokens_backward",
)


class Tuning(ctypes.Structure):
    _fields_ = [("vec", ctypes.c_int), ("staging", ctypes.c_int), ("strip_w", ctypes.c_int),
                ("rows", ctypes.c_int), ("ctas_per_sm", ctypes.c_int), ("force_generic", ctypes.c_int),
                ("order", ctypes.c_int), ("merge", ctypes.c_int), ("geo", ctypes.c_int),
                ("bwd", ctypes.c_int), ("bwd_margin", ctypes.c_int), ("bwd_lanes", ctypes.c_int),
                ("reserved", ctypes.c_int * 4)]


class MSDAError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -m bm2f_b200.build`")
        L = ctypes.CDLL(LIB_PATH)
        vp, i64p, ci = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int
        tp = ctypes.POINTER(Tuning)
        L.bm2f_msda_abi_version.restype = ci
        L.bm2f_msda_build_info.restype = ctypes.c_char_p
        L.bm2f_msda_last_error.restype = ctypes.c_char_p
        L.bm2f_msda_launch_count.restype = ctypes.c_uint64
        L.bm2f_msda_set_default_tuning.argtypes = [tp]
        L.bm2f_msda_set_default_tuning.restype = None
        L.bm2f_msda_debug_phase_profile.argtypes = [vp]
        L.bm2f_msda_debug_phase_profile.restype = None
        L.bm2f_msda_check_im2col_step.argtypes = [ci, ci]
        L.bm2f_msda_check_im2col_step.restype = ci
        L.bm2f_msda_forward.argtypes = [vp, i64p, i64p, vp, vp, vp] + [ci] * 8 + [tp, vp]
        L.bm2f_msda_forward.restype = ci
        L.bm2f_msda_backward.argtypes = [vp, i64p, i64p, vp, vp, vp, vp, vp, vp] + [ci] * 8 + [tp, vp]
        L.bm2f_msda_backward.restype = ci
        L.bm2f_msda_fused_supported.argtypes = [ci] * 5
        L.bm2f_msda_fused_supported.restype = ci
        L.bm2f_msda_fused_forward.argtypes = [vp, i64p, i64p, vp, vp, vp, vp] + [ci] * 8 + [tp, vp]
        L.bm2f_msda_fused_forward.restype = ci
        L.bm2f_msda_fused_backward.argtypes = [vp, i64p, i64p, vp, vp, vp, vp, vp, vp, vp] + [ci] * 8 + [tp, vp]
        L.bm2f_msda_fused_backward.restype = ci
        L.bm2f_linear_workspace_bytes.argtypes = [ci, ci]
        L.bm2f_linear_workspace_bytes.restype = ctypes.c_size_t
        L.bm2f_linear_forward.argtypes = [vp] * 5 + [ci] * 4 + [vp]
        L.bm2f_linear_forward.restype = ci
        L.bm2f_linear_backward_input.argtypes = [vp] * 4 + [ci] * 4 + [vp]
        L.bm2f_linear_backward_input.restype = ci
        L.bm2f_linear_backward_weight.argtypes = [vp] * 4 + [ci] * 4 + [vp]
        L.bm2f_linear_backward_weight.restype = ci
        L.bm2f_msda_forward_backward_host.argtypes = [vp] * 10 + [ci] * 8 + [tp]
        L.bm2f_msda_forward_backward_host.restype = ci
        _lib = L
    return _lib


def last_error() -> str:
    return lib().bm2f_msda_last_error().decode()


def _check(rc: int, what: str):
    if rc != 0:
        raise MSDAError(f"{what} failed ({rc}): {last_error()}")


def make_tuning(**kw):
    t = Tuning()
    for k, v in kw.items():
        if k == "variant":              # A/B selector of the anchor-sorted backward (reserved[0])
            t.reserved[0] = int(v)
        else:
            setattr(t, k, int(v))
    return t


def _tp(t):
    return ctypes.byref(t) if t is not None else None


def forward(value, shapes, start, loc, attn, out, dims, dtype=DTYPE_F32, tuning=None, stream=0):
    """All tensor arguments are device addresses (ints).  dims = (N, S, M, D, L, Lq, P)."""
    _check(lib().bm2f_msda_forward(value, shapes, start, loc, attn, out, *dims, dtype, _tp(tuning), stream),
           "bm2f_msda_forward")


def backward(value, shapes, start, loc, attn, grad_out, grad_value, grad_loc, grad_attn, dims,
             dtype=DTYPE_F32, tuning=None, stream=0):
    _check(lib().bm2f_msda_backward(value, shapes, start, loc, attn, grad_out, grad_value, grad_loc, grad_attn,
                                    *dims, dtype, _tp(tuning), stream), "bm2f_msda_backward")


def forward_backward_host(value, shapes, start, loc, attn, grad_out, out, grad_value, grad_loc, grad_attn, dims,
                          dtype=DTYPE_F32, tuning=None):
    """Host-pointer entry (pinned buffers recommended); grad_out == 0 runs forward only."""
    _check(lib().bm2f_msda_forward_backward_host(value, shapes, start, loc, attn, grad_out, out, grad_value,
                                                 grad_loc, grad_attn, *dims, dtype, _tp(tuning)),
           "bm2f_msda_forward_backward_host")


def fused_forward(value, shapes, start, ref, offsets, logits, out, dims, dtype=DTYPE_F32, tuning=None, stream=0):
    """softmax(logits) and ref + offsets/(W,H) are computed inside the kernel."""
    _check(lib().bm2f_msda_fused_forward(value, shapes, start, ref, offsets, logits, out, *dims, dtype, _tp(tuning),
                                         stream), "bm2f_msda_fused_forward")


def fused_backward(value, shapes, start, ref, offsets, logits, grad_out, grad_value, grad_offsets, grad_logits, dims,
                   dtype=DTYPE_F32, tuning=None, stream=0):
    _check(lib().bm2f_msda_fused_backward(value, shapes, start, ref, offsets, logits, grad_out, grad_value,
                                          grad_offsets, grad_logits, *dims, dtype, _tp(tuning), stream),
           "bm2f_msda_fused_backward")


def linear_forward(x, weight, bias, y, workspace, rows, out_features, in_features, split=3, stream=0):
    """tcgen05 projection GEMM: y = x @ weight^T + bias (device addresses; bias may be 0)."""
    _check(lib().bm2f_linear_forward(x, weight, bias, y, workspace, rows, out_features, in_features, split, stream),
           "bm2f_linear_forward")


def launch_count() -> int:
    return int(lib().bm2f_msda_launch_count())


def set_default_tuning(tuning=None):
    lib().bm2f_msda_set_default_tuning(_tp(tuning))
