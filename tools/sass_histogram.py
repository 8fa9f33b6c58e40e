"""Opcode histogram per kernel of libdsac_b200.so (cuobjdump -sass), so that claims about a kernel's instruction mix
(fp64 vs integer, MUFU count, presence of TMA / tensor-core opcodes) can be checked from a committed file.

    python tools/sass_histogram.py [lib] > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dsac_b200", "libdsac_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print("# cuobjdump -sass opcode histogram per kernel of %s (static instruction counts, sm_100a)" % os.path.basename(lib))
groups = {"fp64": ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"), "fp32": ("FFMA", "FMUL", "FADD", "FSETP", "FFMA2", "FMUL2"),
          "int/logic": ("IMAD", "IADD3", "LOP3", "SHF", "ISETP", "LEA", "PRMT", "SEL", "IABS", "POPC"),
          "memory": ("LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOMG", "ATOMS", "RED", "LDC", "LDCU"),
          "tma/tensor": ("UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "UTCMMA", "TCGEN05")}
for k, c in hist.items():
    tot = sum(c.values())
    print("\nkernel %s: %d instructions" % (k, tot))
    for g, ops in groups.items():
        n = sum(c[o] for o in ops)
        print("  %-10s %6d  (%s)" % (g, n, ", ".join("%s %d" % (o, c[o]) for o in ops if c[o])))
    print("  top: " + ", ".join("%s %d" % (o, n) for o, n in c.most_common(14)))
