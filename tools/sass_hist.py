#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of an object file / shared library (cuobjdump -sass).

    python tools/sass_hist.py dct_b200/build/fwd_quant.cu.o [regex] > profiles/sass_r2_k1.txt

For each kernel whose (demangled) name matches `regex`: total instructions, then opcode counts, with the
tile-movement opcodes (UTMALDG / UTMASTG / UBLKCP / LDGSTS / LDG / STG / LDS / STS / SYNCS) listed first.
"""
import collections
import re
import subprocess
import sys


def main():
    obj = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else re.compile(".")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Za-z0-9_.]+)?)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    names = list(kernels)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
    move = ("UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "LDGSTS", "LDG", "STG", "LDS", "STS", "SYNCS", "LDL", "STL")
    for name, d in zip(names, dem):
        if not pat.search(d):
            continue
        c = kernels[name]
        base = collections.Counter()
        for op, n in c.items():
            base[op.split(".")[0]] += n
        print(f"== {d}\n   {sum(c.values())} instructions")
        print("   tile movement: " + ", ".join(f"{k} {base[k]}" for k in move if base[k]))
        print("   by opcode:     " + ", ".join(f"{k} {v}" for k, v in base.most_common()))
        print("   full:          " + ", ".join(f"{k} {v}" for k, v in c.most_common(40)))


if __name__ == "__main__":
    main()
