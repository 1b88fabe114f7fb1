#!/usr/bin/env python
"""Per-SASS-instruction stall samples from an ncu report (needs --import-source on / -lineinfo):
`python tools/ncu_hot.py report.ncu-rep [--top N] [--range LO HI]` prints the hottest instructions and, with --range,
a contiguous listing (index, samples, dominant stall, SASS)."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    recs = []
    for k, r in enumerate(rows[hi + 1:]):
        if r and r[0] in ("Kernel Name", "Address"):   # next kernel of a multi-kernel report: keep the first only
            break
        if len(r) < len(hdr):
            continue
        n = int(r[ci["# Samples"]] or 0)
        st = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        recs.append((k, n, int(r[ci["Instructions Executed"]] or 0), st, r[ci["Source"]].strip()))
    tot = sum(r[1] for r in recs)
    print(f"total samples {tot}, instructions {len(recs)}")
    if "--range" in sys.argv:
        lo, hi2 = int(sys.argv[sys.argv.index("--range") + 1]), int(sys.argv[sys.argv.index("--range") + 2])
        for k, n, ex, st, src in recs[lo:hi2]:
            print(f"{k:5d} {n:6d} {ex:9d} {st[0][1]:>12s}:{st[0][0]:<5d} {src[:90]}")
        return
    for k, n, ex, st, src in sorted(recs, key=lambda r: -r[1])[:top]:
        print(f"{k:5d} {n:6d} {100*n/tot:5.1f}% {ex:9d} {st[0][1]:>12s}:{st[0][0]:<5d} {st[1][1]:>12s}:{st[1][0]:<5d} {src[:80]}")


if __name__ == "__main__":
    main()
