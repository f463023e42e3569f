#!/usr/bin/env python3
"""Per-kernel SASS summary of libzkmsm.so: registers, stack/spill bytes, instruction mix (which pipe the integer
work lands on).  Output goes to profiles/rNN_sass_summary.txt.

  python tools/sass_summary.py [path/to/lib.so] > profiles/r02_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "zkvm_b200/libzkmsm.so"
txt = subprocess.run(["cuobjdump", "-sass", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout

res = {}
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        res[cur] = tuple(int(x) for x in m.groups()); cur = None

mix = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); mix[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        mix[cur][m.group(1)] += 1


def demangle(n):
    m = re.search(r"\d+(k_[a-z0-9_]+?)E", n)
    return m.group(1) if m else n


GROUPS = [
    ("IMAD.WIDE*", lambda o: o.startswith("IMAD.WIDE")),
    ("IMAD.HI*", lambda o: o.startswith("IMAD.HI")),
    ("IMAD (32-bit mul)", lambda o: o in ("IMAD", "IMAD.U32", "IMAD.SHL.U32") or o.startswith("IMAD.SHL")),
    ("IMAD.X/IADD (adds on the FMA pipe)", lambda o: o.startswith("IMAD.X") or o.startswith("IMAD.IADD")),
    ("IMAD.MOV (moves on the FMA pipe)", lambda o: o.startswith("IMAD.MOV")),
    ("IADD3*/IADD* (ALU pipe)", lambda o: o.startswith("IADD")),
    ("LOP3/SHF/SEL/ISETP/MOV/PRMT", lambda o: o.split(".")[0] in ("LOP3", "SHF", "SEL", "ISETP", "MOV", "PRMT", "PLOP3", "LEA")),
    ("SHFL", lambda o: o.startswith("SHFL")),
    ("LDG/STG/LD/ST", lambda o: o.split(".")[0] in ("LDG", "STG", "LD", "ST")),
    ("LDL/STL (local = spills/arrays)", lambda o: o.split(".")[0] in ("LDL", "STL")),
    ("LDS/STS/ATOMS", lambda o: o.split(".")[0] in ("LDS", "STS", "ATOMS")),
    ("ATOMG/RED", lambda o: o.split(".")[0] in ("ATOMG", "RED", "ATOM")),
    ("BRA/BSSY/BSYNC/CALL/RET", lambda o: o.split(".")[0] in ("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "WARPSYNC", "BAR")),
]

print(f"# SASS summary of {LIB} (cuobjdump -sass -res-usage), static instruction counts per kernel")
print("# wgmma/tcgen05 are absent by design: the path is integer limb arithmetic on the IMAD pipe (DESIGN.md section 4)")
for fn, c in mix.items():
    name = demangle(fn)
    r = res.get(fn)
    total = sum(c.values())
    print(f"\n== {name}: {total} instructions" + (f", REG {r[0]}, STACK {r[1]} B, SHARED {r[2]} B, LOCAL {r[3]} B" if r else ""))
    covered = 0
    for label, pred in GROUPS:
        k = sum(v for o, v in c.items() if pred(o))
        covered += k
        if k:
            print(f"   {label:42s} {k:7d}  ({100.0 * k / total:5.1f} %)")
    print(f"   {'other':42s} {total - covered:7d}")
    top = ", ".join(f"{o} {v}" for o, v in c.most_common(8))
    print(f"   top opcodes: {top}")
