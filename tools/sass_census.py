"""SASS census of the shipped default kernels (runs without a GPU): python tools/sass_census.py > profiles/rNN_sass_main_kernels.txt
For every kernel whose mangled name matches one of the patterns below: instruction count, mnemonic census, the Blackwell
data-movement / tensor-core mnemonics present (UTMALDG = TMA load, UTMASTG = TMA store, UTCHMMA = tcgen05.mma, LDTM =
tcgen05.ld, REDG = red.global, SYNCS = mbarrier), then the full listing."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bm2f_b200", "libbm2f_msda.so")
# (label, regex on the mangled name): the kernels the default paths launch at the benchmark shapes
KERNELS = [
    ("forward, default: geometry warps + 16-byte records (fp32, L = 3, 28 consumer + 3 geometry warps)",
     r"msda_fwd_geo_kernelIfLi3ELi4ELi32ELi28ELi3ELi4ELb0ELb0ELi1ELi2E"),
    ("backward, default for batch x queries >= 65536: anchor-sorted (L = 3, 8 warps, 4 lanes per point, 2 CTAs / SM)",
     r"msda_bwd_sorted_kernelILi3ELi6ELi8ELi4ELb0ELi8192ELi2ELi0E"),
    ("backward, default for small launches: per-corner REDs (fp32, L = 3)", r"msda_bwd_fast_kernelIfLi4ELi3ELi4ELi32ELi16ELi1ELb1ELi1ELb0ELb0E"),
    ("fused forward (module path): softmax + locations in a prologue warp", r"msda_fwd_fast_kernelIfLi4ELi3ELi4ELi32ELi16ELi1ELb1ELi1ELb1E"),
    ("projection GEMM, tf32x3, persistent, N = 256", r"linear_tf32x3_persistent_kernelILi256ELi1ELi4ELi3ELb0ELb0ELb0ELi1E"),
    ("projection GEMM, single TF32 pass, TMA activations, N = 256", r"linear_tf32x3_persistent_kernelILi256ELi1ELi4ELi3ELb0ELb1ELb0ELi1E"),
    ("3x3 convolution, single TF32 pass, TMA row-shifted activations, two row tiles per weight block", r"linear_tf32x3_persistent_kernelILi256ELi1ELi4ELi3ELb0ELb1ELb1ELi2E"),
    ("weight gradient, tf32x3, transposing producers", r"linear_dw_tf32x3_kernel"),
    ("weight gradient, single TF32 pass, TMA-fed MN-major operands", r"linear_dw_tma_kernel"),
]
MARK = ("UTMALDG", "UTMASTG", "UTMAREDG", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "REDG", "RED.", "SYNCS", "UBLKCP", "LDG.E.128", "LDS.128", "ATOMS", "FFMA2")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs.setdefault(cur, [])            # static kernels appear once per translation unit: keep the first body
            if funcs[cur]:
                cur = None
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append(m.group(1).strip())
    print("# cuobjdump -sass bm2f_b200/libbm2f_msda.so — kernels launched by the default paths (tools/sass_census.py)\n")
    full = "--full" in sys.argv
    for label, pat in KERNELS:
        names = [n for n in funcs if re.search(pat, n) and funcs[n]]
        if not names:
            print(f"==== {label}\n# no kernel matches {pat}\n")
            continue
        n = names[0]
        ins = funcs[n]
        census = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0] for i in ins)
        marks = {k: sum(v for m_, v in census.items() if k in m_) for k in MARK}
        print(f"==== {label}\n# {n}")
        print(f"# {len(ins)} SASS instructions; " + ", ".join(f"{k} {v}" for k, v in marks.items() if v))
        print("# census: " + ", ".join(f"{k} {v}" for k, v in census.most_common(28)))
        if full:
            print("\n".join(ins))
        print()


if __name__ == "__main__":
    main()
