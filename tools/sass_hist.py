#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel in an object file / library (no GPU needed).
Usage: tools/sass_hist.py <file.o|.so> <kernel-name-substring> [--dump out.txt]
Splits the kernel at backward branches to show the loops (innermost bodies), so instruction counts per loop trip can be read
before spending GPU time."""
import re
import subprocess
import sys
from collections import Counter


def main():
    path, pat = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    sel = [f for f in funcs[1:] if pat in f.split("\n", 1)[0]]
    if not sel:
        sys.exit("no kernel matches " + pat)
    for f in sel:
        name = f.split("\n", 1)[0]
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        if "--dump" in sys.argv:
            open(sys.argv[sys.argv.index("--dump") + 1], "w").write("\n".join(f"{a:06x} {s}" for a, s in ins))
        def op(s):
            s = re.sub(r"^@!?U?P\w+\s+", "", s)
            return s.split()[0].split(".")[0]
        tot = Counter(op(s) for _, s in ins)
        print(f"== {name[:110]}\n   {len(ins)} instructions; top: " + ", ".join(f"{k} {v}" for k, v in tot.most_common(12)))
        # loops = backward branches
        loops = []
        for a, s in ins:
            m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?(\w+)\)?", s)
            m2 = re.search(r"BRA.*0x([0-9a-f]+)", s)
            if m2:
                tgt = int(m2.group(1), 16)
                if tgt < a:
                    loops.append((tgt, a))
        for tgt, a in sorted(loops, key=lambda x: x[1] - x[0]):
            body = [s for aa, s in ins if tgt <= aa <= a]
            c = Counter(op(s) for s in body)
            print(f"   loop {tgt:06x}..{a:06x}: {len(body)} instr; " + ", ".join(f"{k} {v}" for k, v in c.most_common(14)))


if __name__ == "__main__":
    main()
