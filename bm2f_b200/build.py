"""In-tree build of the native pieces (no JIT cache: the .so files travel with the tree).

  libbm2f_msda.so                     nvcc, sm_100a only: kernels + C ABI (include/bm2f_msda.h)
  MultiScaleDeformableAttention*.so   g++: torch-facing module with the reference's two functions
  msda_microbench                     nvcc: stand-alone L1/L2 gather + vector-RED microbenchmarks

Usage:  python -m bm2f_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libbm2f_msda.so")
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
EXT = os.path.join(PKG, "MultiScaleDeformableAttention" + EXT_SUFFIX)
MICRO = os.path.join(PKG, "msda_microbench")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _gxx() -> str:
    # /opt/gcc/bin wrappers exist in this image; prefer the system compiler
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-8000:])
        raise RuntimeError("build command failed: " + " ".join(cmd))
    return r.stdout


# translation units of libbm2f_msda.so (compiled in parallel, then linked)
LIB_UNITS = ("api_common.cu", "msda_api.cu", "msda_bwd_sorted.cu", "linear_api.cu", "glue_api.cu", "fpn_api.cu", "host_api.cu")
OBJ_DIR = os.path.join(PKG, "_obj")


def build_lib(force=False, ptxas_log=None) -> str:
    """nvcc -c every unit whose sources changed (headers are shared: any .cuh change rebuilds all), then link.
    BM2F_SWEEP=1 in the environment adds the sweep-only kernel instantiations (tools/sweep.py)."""
    import glob
    from concurrent.futures import ThreadPoolExecutor
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(ROOT, "include", "bm2f_msda.h")]
    flags = list(NVCC_FLAGS)
    sweep = os.environ.get("BM2F_SWEEP", "0") not in ("", "0")
    if sweep:
        flags.append("-DBM2F_SWEEP")
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "flags.txt")
    if not os.path.exists(stamp) or open(stamp).read() != " ".join(flags):
        force = True
    units = [u for u in LIB_UNITS if os.path.exists(os.path.join(CSRC, u))]
    todo = []
    for u in units:
        obj = os.path.join(OBJ_DIR, u[:-3] + ".o")
        if force or _stale(obj, headers + [os.path.join(CSRC, u)]):
            todo.append((u, obj))

    def compile_unit(job):
        u, obj = job
        return u, _run([_nvcc()] + flags + ["-Xptxas", "-v", "-c", "-o", obj, os.path.join(CSRC, u)])

    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            logs = list(ex.map(compile_unit, todo))
        log_path = ptxas_log or os.path.join(OBJ_DIR, "ptxas_v.log")      # build artefact, git-ignored
        with open(log_path, "a" if len(todo) < len(units) else "w") as f:
            for u, out in logs:
                f.write(f"==== {u}\n{out}\n")
        with open(stamp, "w") as f:
            f.write(" ".join(flags))
    objs = [os.path.join(OBJ_DIR, u[:-3] + ".o") for u in units]
    if todo or _stale(LIB, objs):
        _run([_nvcc()] + NVCC_FLAGS + ["-shared", "-o", LIB] + objs)
    return LIB


def build_microbench(force=False) -> str:
    src = os.path.join(CSRC, "msda_microbench.cu")
    if os.path.exists(src) and (force or _stale(MICRO, [src])):
        _run([_nvcc()] + NVCC_FLAGS + ["-o", MICRO, src])
    return MICRO


def build_ext(force=False) -> str:
    src = os.path.join(CSRC, "msda_torch_ext.cpp")
    if not (force or _stale(EXT, [src, LIB, os.path.join(ROOT, "include", "bm2f_msda.h")])):
        return EXT
    import torch
    from torch.utils import cpp_extension as ce

    inc = ce.include_paths("cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths()
    inc = list(inc) + [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
    libdirs = ce.library_paths()
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = [_gxx(), "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-attributes",
           "-DTORCH_EXTENSION_NAME=MultiScaleDeformableAttention", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={abi}", src]
    for i in inc:
        cmd += ["-isystem", i]
    for d in libdirs:
        cmd += ["-L" + d, "-Wl,-rpath," + d]
    cmd += ["-L" + PKG, "-lbm2f_msda", "-Wl,-rpath,$ORIGIN",
            "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
            "-o", EXT]
    _run(cmd)
    return EXT


def build_all(force=False):
    build_lib(force)
    build_ext(force)
    build_microbench(force)
    return LIB, EXT


if __name__ == "__main__":
    out = build_all(force="--force" in sys.argv)
    print("\n".join(out))
