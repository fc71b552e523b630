#!/usr/bin/env python
"""List the loops (backward branches) of every kernel whose name matches a pattern, with their instruction mix.
    nvcc ... -cubin -o x.cubin file.cu && python tools/sass_loops.py x.cubin k_prep_stream
"""
import re
import subprocess
import sys
from collections import Counter


def main(cubin, pat):
    txt = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n")[0]
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if not re.search(pat, dem):
            continue
        ins = []
        for l in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        print(re.sub(r"\(anonymous namespace\)::|akz::", "", dem)[:90], "|", len(ins), "instructions")
        for a, t in ins:
            m = re.search(r"BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                tgt = int(m.group(1), 16)
                c = Counter()
                for _, b in [x for x in ins if tgt <= x[0] <= a]:
                    parts = b.split()
                    op = parts[1] if parts[0].startswith("@") else parts[0]
                    c[op.split(".")[0]] += 1
                print("   loop %05x..%05x n=%d " % (tgt, a, sum(c.values())), dict(c.most_common(16)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ".")
