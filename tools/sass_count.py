#!/usr/bin/env python
"""tools/sass_count.py [pattern] -- static SASS opcode histogram per kernel of libothello_b200.so"""
import collections
import re
import subprocess
import sys

so = "subproc_b200/libothello_b200.so"
pat = sys.argv[1] if len(sys.argv) > 1 else "playout_kernel"
txt = subprocess.run("cuobjdump -sass %s | c++filt" % so, shell=True, capture_output=True, text=True).stdout
ALU = {"LOP3", "SHF", "SEL", "ISETP", "IADD3", "LEA", "VIADD", "PRMT", "IADD", "MOV", "IMNMX", "VIMNMX", "PLOP3", "ICMP", "BMSK", "SGXT", "IABS", "FSEL", "FSETP", "FMNMX"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD"}
XU = {"POPC", "BREV", "FLO", "I2F", "F2I", "I2FP", "MUFU"}
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if pat not in name:
        continue
    ops = collections.Counter()
    for line in f.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(2).split(".")[0]] += 1
    n = sum(ops.values())
    alu = sum(v for k, v in ops.items() if k in ALU)
    fma = sum(v for k, v in ops.items() if k in FMA)
    xu = sum(v for k, v in ops.items() if k in XU)
    print("%s\n   total %d  alu %d  fma %d  xu %d  other %d" % (name[:110], n, alu, fma, xu, n - alu - fma - xu))
    print("   ", dict(ops.most_common(16)))
