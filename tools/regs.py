"""Summarise `nvcc -Xptxas -v` output: registers / spills / smem per kernel instantiation."""
import re, subprocess, sys
log = open(sys.argv[1]).read()
pat = (r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
       r"(\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers, used (\d+) barriers(?:, (\d+) bytes smem)?")
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for name, stack, ss, sl, regs, bars, smem in re.findall(pat, log):
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    d = d.replace("bm2f::", "").replace("(FastParams, CUtensorMap_st, CUtensorMap_st)", "").replace("void ", "")
    if flt in d:
        print(f"{regs:>4} regs  stack {stack:>4}  spill {ss:>4}/{sl:<4}  smem {smem or 0:>6}  {d[:100]}")
