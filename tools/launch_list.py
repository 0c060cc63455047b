#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum[,...] --csv --log-file X <command>`) per kernel.

usage: tools/launch_list.py X.csv [--md]     prints kernel | launches | avg ms | total ms | share of the listed GPU time
The times are cold-cache and serialised (one kernel at a time under the profiler): the SHARES are what is comparable with the
bench's own CUDA-event times, not the absolutes."""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    md = "--md" in sys.argv
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    head = rows[0]
    ki, mi, vi, ui, ii = (head.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = OrderedDict()
    extra = {}
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("fdb::", "")
        val = float(r[vi].replace(",", "")) if r[vi] not in ("", "n/a") else 0.0
        if r[mi] == "gpu__time_duration.sum":
            scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[ui], 1e-6)
            per.setdefault(name, []).append(val * scale)
        else:
            extra.setdefault(name, {}).setdefault(r[mi], []).append(val)
    total = sum(sum(v) for v in per.values())
    cols = sorted({m for e in extra.values() for m in e})
    hdr = ["kernel", "launches", "avg ms", "total ms", "share %"] + [c.split("__")[-1] for c in cols]
    if md:
        print("| " + " | ".join(hdr) + " |")
        print("|" + "---|" * len(hdr))
    else:
        print("\t".join(hdr))
    for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        cells = [f"`{name[:90]}`" if md else name[:90], str(len(v)), f"{sum(v) / len(v):.4f}", f"{sum(v):.3f}", f"{100 * sum(v) / total:.2f}"]
        for c in cols:
            xs = extra.get(name, {}).get(c, [])
            cells.append(f"{sum(xs) / len(xs):.4g}" if xs else "")
        print(("| " + " | ".join(cells) + " |") if md else "\t".join(cells))
    print(f"\ntotal listed GPU time: {total:.3f} ms over {sum(len(v) for v in per.values())} launches")


if __name__ == "__main__":
    main()
