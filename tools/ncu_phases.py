"""Aggregate an ncu --import-source report by source-line ranges of one file:
   python tools/ncu_phases.py report.ncu-rep file.cuh name:lo-hi name:lo-hi ...   (other files are listed by file)"""
import csv
import subprocess
import sys


def main():
    rep, fname = sys.argv[1], sys.argv[2]
    ranges = []
    for a in sys.argv[3:]:
        n, r = a.split(":")
        lo, hi = r.split("-")
        ranges.append((n, int(lo), int(hi)))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr = None, None
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r[0].strip().isdigit():
            continue
        nm = len(hdr) - 4
        d = dict(zip(hdr[4:], r[-nm:]))
        ln = int(r[0])
        key = cur
        if cur == fname:
            key = "%s:other" % fname
            for n, lo, hi in ranges:
                if lo <= ln <= hi:
                    key = n
                    break
        def num(k):
            try:
                return int(d.get(k, "0") or 0)
            except ValueError:
                return 0
        a = agg.setdefault(key, [0, 0])
        a[0] += num("Instructions Executed")
        a[1] += num("# Samples")
    ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
    print(f"# total warp instructions {ti}, samples {ts}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"{k:32s} inst {a[0]:12d} {100.0 * a[0] / ti:5.1f}%   samples {100.0 * a[1] / max(ts, 1):5.1f}%")


if __name__ == "__main__":
    main()
