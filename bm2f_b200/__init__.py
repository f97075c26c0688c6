"""bm2f_b200 — B200-native multi-scale deformable attention for Mask2Former / BM2F.

The package holds only what the MSDeformAttn hot path needs:

  csrc/                          sm_100a kernels + the C ABI (include/bm2f_msda.h)
  libbm2f_msda.so                built from csrc/ by bm2f_b200.build
  MultiScaleDeformableAttention  torch extension with the reference's two entry points
                                 (ops/src/vision.cpp:18-21) — the drop-in boundary
  cabi                           ctypes view of the same C ABI (tests, bench, non-torch callers)
  ops/                           mirror of the reference's ops/functions + ops/modules
  workloads                      synthetic Mask2Former-shaped inputs

There is no CPU implementation and no fallback: importing the extension fails loudly when the
native library is missing, and every op raises on non-CUDA tensors.
"""
from __future__ import annotations

import importlib.util
import os
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_EXT_NAME = "MultiScaleDeformableAttention"


def load_extension():
    """Import the in-tree torch extension and register it under the reference's module name,
    so `import MultiScaleDeformableAttention as MSDA` (ops/functions/ms_deform_attn_func.py:21)
    resolves to the B200 build."""
    if _EXT_NAME in sys.modules and getattr(sys.modules[_EXT_NAME], "ms_deform_attn_forward", None):
        return sys.modules[_EXT_NAME]
    import sysconfig

    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    path = os.path.join(_PKG, _EXT_NAME + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build the native library first with `python -m bm2f_b200.build` "
            "(nvcc for sm_100a + g++); there is no Python or CPU fallback for this op.")
    spec = importlib.util.spec_from_file_location(_EXT_NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[_EXT_NAME] = mod
    return mod


__all__ = ["load_extension"]
