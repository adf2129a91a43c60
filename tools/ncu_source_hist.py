#!/usr/bin/env python
"""Opcode histogram + stall hot-spots of one kernel from `ncu -i rep --page source --csv --kernel-name regex:<k>`.
Usage: tools/ncu_source_hist.py src.csv <warp_steps>   (warp_steps = samples x T: prints instructions per warp-step)."""
import csv
import re
import sys
from collections import Counter, defaultdict


def main(path, units):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    ops, stall_by_op = Counter(), defaultdict(Counter)
    tot = 0
    lines = []
    for r in rows[2:]:
        if len(r) < len(hdr) or r[0] == hdr[0] or not r[col["Instructions Executed"]].isdigit():
            continue
        src = r[col["Source"]].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        op = m.group(2) if m else src
        base = op.split(".")[0]
        if base in ("LDG", "STG", "LDS", "STS", "LDGSTS"):
            base = ".".join(op.split(".")[:1] + [p for p in op.split(".")[1:] if p in ("64", "128", "E")][:2])
        n = int(r[col["Instructions Executed"]] or 0)
        ops[base] += n
        tot += n
        samples = int(r[col["# Samples"]] or 0)
        st = {k[6:]: int(r[col[k]] or 0) for k in col if k.startswith("stall_") and "(" not in k}
        lines.append((samples, n, src, st))
    print(f"total warp instructions {tot}  per unit {tot / units:.1f}")
    print("| opcode | per unit | share |\n|---|---:|---:|")
    for k, v in ops.most_common(28):
        print(f"| {k} | {v / units:.1f} | {100 * v / tot:.1f}% |")
    tots = Counter()
    for s, n, src, st in lines:
        for k, v in st.items():
            tots[k] += v
    allsamp = sum(s for s, *_ in lines)
    print("\nstall samples by reason:", {k: round(v / max(1, allsamp), 3) for k, v in tots.most_common(8)})
    print("\ntop stall sites:")
    for s, n, src, st in sorted(lines, key=lambda x: -x[0])[:25]:
        top = ", ".join(f"{k}={v}" for k, v in Counter(st).most_common(2) if v)
        print(f"  {s:6d} ({100 * s / allsamp:4.1f}%)  x{n / units:5.2f}  {src[:70]:70s} {top}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]))
