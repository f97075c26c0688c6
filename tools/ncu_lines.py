"""Per-source-line summary of an ncu report captured with --import-source on:
   python tools/ncu_lines.py report.ncu-rep [min_pct]
prints, for every CUDA source line above min_pct of the kernel's executed warp instructions or stall samples, its share of
both and the dominant stall reasons.  (ncu -i ... --page source --print-source cuda,sass --csv, aggregated.)"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    files = {}
    cur_file, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "Function Name":
            continue
        if hdr is None or not r[0].strip().isdigit():
            continue
        nm = len(hdr) - 4                     # metric columns; source text with quotes can break the left columns
        d = dict(zip(hdr[4:], r[-nm:]))
        files.setdefault(cur_file, []).append((int(r[0]), r[1], d))
    stall_keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    def num(d, k):
        try:
            return int(d.get(k, "0") or 0)
        except ValueError:
            return 0
    tot_i = sum(num(d, "Instructions Executed") for f in files.values() for _, _, d in f)
    tot_s = sum(num(d, "# Samples") for f in files.values() for _, _, d in f)
    print(f"# total warp instructions {tot_i}, stall samples {tot_s}")
    agg = {}
    for f, lines in files.items():
        for ln, src, d in lines:
            n = num(d, "Instructions Executed")
            s = num(d, "# Samples")
            for k in stall_keys:
                agg[k] = agg.get(k, 0) + num(d, k)
            if 100.0 * n / max(tot_i, 1) >= min_pct or 100.0 * s / max(tot_s, 1) >= min_pct:
                top = sorted(((num(d, k), k[6:]) for k in stall_keys), reverse=True)[:3]
                tops = " ".join(f"{k}:{100.0 * v / max(s, 1):.0f}%" for v, k in top if v)
                print(f"{f[:22]:22s} {ln:4d} inst {100.0 * n / tot_i:5.1f}%  samples {100.0 * s / max(tot_s, 1):5.1f}%  [{tops}]  {src.strip()[:90]}")
    print("# stall reasons overall: " + " ".join(f"{k[6:]}:{100.0 * v / max(tot_s, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))


if __name__ == "__main__":
    main()
