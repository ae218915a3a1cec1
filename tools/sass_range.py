#!/usr/bin/env python
"""tools/sass_range.py KERNEL_SUBSTR START END [START END ...] -- opcode histogram over address ranges"""
import collections, re, subprocess, sys
so = "subproc_b200/libothello_b200.so"
txt = subprocess.run("cuobjdump -sass %s | c++filt" % so, shell=True, capture_output=True, text=True).stdout
pat = sys.argv[1]
rng = [(int(sys.argv[i], 16), int(sys.argv[i + 1], 16)) for i in range(2, len(sys.argv) - 1, 2)]
show = "--show" in sys.argv
ALU = {"LOP3", "SHF", "SEL", "ISETP", "IADD3", "LEA", "VIADD", "PRMT", "IADD", "MOV", "PLOP3", "IMNMX", "VIMNMX"}
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    if pat not in f.split("\n")[0]:
        continue
    ops = collections.Counter()
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
        if not m:
            continue
        a = int(m.group(1), 16)
        if any(lo <= a < hi for lo, hi in rng):
            op = m.group(3)
            ops[op if show else op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("IMAD") and "." in op else "")] += 1
            if show:
                print("%04x %s %s%s" % (a, m.group(2) or "", op, m.group(4)))
    n = sum(ops.values())
    alu = sum(v for k, v in ops.items() if k.split(".")[0] in ALU)
    print(f.split("\n")[0][:100])
    print("  total", n, "alu", alu, dict(ops.most_common(30)))
    break
