"""TEST / BASELINE INFRASTRUCTURE — builds the reference's own CUDA op into oracle/_ref/.

The reference (Deformable-DETR's MultiScaleDeformableAttention op as bundled by BM2F) is compiled
from its sources IN PLACE under /root/reference — nothing is copied into the repository — for
sm_100a, under the module name `msda_reference_cuda` so it can be imported next to the B200
build.  Output: oracle/_ref/msda_reference_cuda<EXT_SUFFIX> (git-ignored, travels to the GPU box).

It is used only (a) by bench.py to time "the reference's own CUDA op" beside the new kernels and
(b) by tests/test_gpu_reference_op.py to record its parity against the same oracle.
stage_ops_py() additionally stages the reference's Python files of the op package (test.py, functions/, modules/)
unmodified into oracle/_ref/ops_unmodified/ for tests/test_gpu_reference_dropin.py.

Run: python oracle/build_ref.py     (needs /root/reference; a no-op with a message otherwise)
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/mask2former/modeling/pixel_decoder/ops/src"
OUT_DIR = os.path.join(HERE, "_ref")
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
OUT = os.path.join(OUT_DIR, "msda_reference_cuda" + EXT_SUFFIX)
NAME = "msda_reference_cuda"


def build(force=False):
    if not os.path.isdir(REF_SRC):
        print("reference sources not present (%s): keeping whatever is in oracle/_ref/" % REF_SRC)
        return OUT if os.path.exists(OUT) else None
    if os.path.exists(OUT) and not force:
        return OUT
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT_DIR, exist_ok=True)
    inc = list(ce.include_paths()) + [sysconfig.get_paths()["include"], "/usr/local/cuda/include", REF_SRC]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = ["-DWITH_CUDA", f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
              f"-D_GLIBCXX_USE_CXX11_ABI={abi}"]
    incflags = []
    for i in inc:
        incflags += ["-isystem", i] if i != REF_SRC else ["-I", i]
    objs = []
    nvcc = "/usr/local/cuda/bin/nvcc"
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cu_obj = os.path.join(OUT_DIR, "ref_cuda_unit.o")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-w", "-Xcompiler", "-w", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
           "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__", "-ccbin", gxx,
           "-c", os.path.join(HERE, "ref_wrap", "ref_cuda_unit.cu"), "-o", cu_obj] + common + incflags
    print("nvcc reference .cu (about 2 min) ...", flush=True)
    subprocess.check_call(cmd)
    objs.append(cu_obj)
    for src in ("vision.cpp", os.path.join("cpu", "ms_deform_attn_cpu.cpp")):
        o = os.path.join(OUT_DIR, os.path.basename(src) + ".o")
        subprocess.check_call([gxx, "-O2", "-fPIC", "-std=c++17", "-w", "-c",
                               os.path.join(REF_SRC, src), "-o", o] + common + incflags)
        objs.append(o)
    link = [gxx, "-shared", "-o", OUT] + objs
    for d in ce.library_paths("cuda") if "device_type" in ce.library_paths.__code__.co_varnames else ce.library_paths():
        link += ["-L" + d, "-Wl,-rpath," + d]
    link += ["-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
             "-ltorch_python", "-lcudart"]
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    return OUT


REF_OPS = "/root/reference/mask2former/modeling/pixel_decoder/ops"
STAGE_DIR = os.path.join(OUT_DIR, "ops_unmodified")
STAGED = ("test.py", "functions/__init__.py", "functions/ms_deform_attn_func.py", "modules/__init__.py",
          "modules/ms_deform_attn.py")
MANIFEST = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_ops_sha256.json")


def stage_ops_py():
    """Stage the reference's own Python files of the op package — its only test (ops/test.py), its autograd function
    and its nn.Module — byte for byte into oracle/_ref/ops_unmodified/ops/ (git-ignored, travels to the GPU box), so
    that tests/test_gpu_reference_dropin.py can run them UNCHANGED on top of this repository's
    `MultiScaleDeformableAttention` extension ("test.py must run unchanged against the new
    module").  Nothing is copied into the tracked tree; the committed manifest holds only SHA-256 digests, which the
    test uses to prove the staged files are the reference's."""
    import hashlib
    import json
    import shutil
    if not os.path.isdir(REF_OPS):
        return STAGE_DIR if os.path.isdir(STAGE_DIR) else None
    digests = {}
    for rel in STAGED:
        dst = os.path.join(STAGE_DIR, "ops", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_OPS, rel), dst)
        digests[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    if not os.path.exists(MANIFEST) or json.load(open(MANIFEST)) != digests:
        with open(MANIFEST, "w") as f:
            json.dump(digests, f, indent=1, sort_keys=True)
            f.write("\n")
    return STAGE_DIR


def load():
    """Import oracle/_ref/msda_reference_cuda (None when it was never built)."""
    import importlib.util

    import torch  # noqa: F401
    if not os.path.exists(OUT):
        return None
    spec = importlib.util.spec_from_file_location(NAME, OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(stage_ops_py())
