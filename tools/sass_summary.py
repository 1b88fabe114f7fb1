#!/usr/bin/env python
"""Per-kernel SASS evidence of the shipped library: counts of the Blackwell tensor / TMA / TMEM mnemonics (UTCHMMA = tcgen05.mma,
UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
SYNCS = mbarrier), of the warp-level HMMA (mma.sync) and of the CUDA-core pipes that bound the non-GEMM kernels (MUFU, FFMA2).
    python tools/sass_summary.py [lib.so] > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fbanet_b200", "csrc", "libfbanet_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "MUFU.TANH", "MUFU.EX2", "FFMA2", "FHFMA", "BAR.SYNC", "LDL", "STL"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = re.sub(r"\(.*$", "", name).replace("fbanet::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[name]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[name][k] += 1
print(f"# SASS mnemonic counts per kernel of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a); kernels with tensor-core / TMA instructions first")
print(f"{'kernel':72s} {'instr':>6s} " + " ".join(f"{k:>9s}" for k in KEYS))
rows = sorted(counts.items(), key=lambda kv: -(kv[1]["UTCHMMA"] * 1000 + kv[1]["HMMA"] * 10 + kv[1]["UTMALDG"]))
for n, c in rows:
    print(f"{n[:72]:72s} {c['_total']:6d} " + " ".join(f"{c[k]:9d}" for k in KEYS))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"{'TOTAL':72s} {tot['_total']:6d} " + " ".join(f"{tot[k]:9d}" for k in KEYS))
