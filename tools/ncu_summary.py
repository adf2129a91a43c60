#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (per kernel: launches,
total ms, share).  Usage: tools/ncu_summary.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[0] == "ID":
            continue
        name = re.sub(r"lqgk::", "", r[ki])
        name = re.sub(r"\(.*", "", name)
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] == "ns" else (v / 1e3 if r[ui] in ("us", "usecond") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if ms / tot < 5e-4:
            continue
        print(f"| `{k[:90]}` | {n} | {ms:.3f} | {100 * ms / tot:.1f}% |")
    print(f"| **total** | | {tot:.3f} | |")


if __name__ == "__main__":
    main(sys.argv[1])
