#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`) of
`bench.py --steps K --no-graph`: per kernel family the launches, serialized time and DRAM bytes of ONE forward (a forward starts at the head
conv; the last complete run of `launches_per_step` launches is used), written as JSON for bench.py / DESIGN.md.

    python tools/launch_summary.py gpurun_out/launches.csv LAUNCHES_PER_STEP profiles/rN_kernel_summary.json
"""
import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict


def family(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("fbanet::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"<.*$", "", name)
    return name.strip()


def main():
    path, per_step, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hi = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr = rows[hi]
    ci = {h: i for i, h in enumerate(hdr)}
    launches = OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[0].isdigit():
            continue
        d = launches.setdefault(int(r[ci["ID"]]), {"name": r[ci["Kernel Name"]]})
        val = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        m = r[ci["Metric Name"]]
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3,
                 "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[m] = val * scale
    # this repo's kernels: everything that is not a torch / cub / cuBLAS kernel (a capture taken with `-k regex:fbanet --kernel-name-base
    # demangled` lists base names without the namespace)
    foreign = ("at::", "cub::", "cublas", "cutlass", "thrust::", "elementwise_kernel", "vectorized_", "nccl", "distribution_", "reduce_kernel", "index_")
    ours = [d for d in launches.values() if not any(f in d["name"] for f in foreign)]
    # a forward starts with the head conv: take the last COMPLETE forward of the capture
    starts = [i for i, d in enumerate(ours) if "head_conv" in d["name"]]
    segs = [(a, b) for a, b in zip(starts, starts[1:]) if b - a == per_step]
    assert segs, f"no complete forward of {per_step} launches between head-conv launches (gaps: {[b - a for a, b in zip(starts, starts[1:])]})"
    step = ours[segs[-1][0]:segs[-1][1]]
    fam = defaultdict(lambda: {"launches": 0, "ms": 0.0, "dram_read": 0.0, "dram_write": 0.0})
    for d in step:
        f = fam[family(d["name"])]
        f["launches"] += 1
        f["ms"] += d.get("gpu__time_duration.sum", 0.0)
        f["dram_read"] += d.get("dram__bytes_read.sum", 0.0)
        f["dram_write"] += d.get("dram__bytes_write.sum", 0.0)
    # the implicit-GEMM launches in launch order (bench.py matches them with its own launch sequence to split DRAM traffic by class)
    conv_order = [{"ms": d.get("gpu__time_duration.sum", 0.0), "dram_bytes": d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)}
                  for d in step if "conv_gemm" in d["name"]]
    tot = sum(f["ms"] for f in fam.values())
    res = {"source": path, "launches_per_step": len(step), "serialized_ms": tot, "families": {}, "conv_gemm_launches": conv_order}
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        f["share"] = f["ms"] / tot
        f["dram_bytes"] = f["dram_read"] + f["dram_write"]
        res["families"][k] = f
        print(f"{k:44s} n={f['launches']:4d} {f['ms']:8.3f} ms {100 * f['share']:5.1f}%  dram {f['dram_bytes'] / 1e9:7.2f} GB")
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
