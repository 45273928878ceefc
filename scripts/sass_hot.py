#!/usr/bin/env python
"""Loop census of one kernel's SASS: sass_hot.py file.o mangled_substring [min_instructions]."""
import re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
minn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, ops = None, []
for l in txt.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1); continue
    if cur and pat in cur:
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if m: ops.append((int(m.group(1), 16), m.group(3), l))
loops = []
for a, op, l in ops:
    if op.startswith("BRA"):
        m = re.search(r"(0x[0-9a-f]+) ;", l)
        if m and int(m.group(1), 16) < a: loops.append((int(m.group(1), 16), a))
print("instructions", len(ops))
for t, a in loops:
    body = [o for o in ops if t <= o[0] <= a]
    if len(body) < minn or len(body) > 3000: continue
    c = lambda p: sum(1 for o in body if re.match(p, o[1]))
    print(f"{t:#x}-{a:#x} n={len(body)} STL={c('STL')} LDL={c('LDL')} LDG={c('LDG')} LDS={c('LDS')} STS={c('STS')} FFMA2={c('FFMA2')} "
          f"FFMA={c('FFMA$|FFMA[.]')} FMUL={c('FMUL')} FADD={c('FADD')} ATOMS={c('ATOMS')} SYNCS={c('SYNCS')} MUFU={c('MUFU')} I={c('IMAD|IADD|LEA|LOP|SHF|ISETP|VIADD|MOV|SEL|PRMT')}")
