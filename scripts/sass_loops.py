#!/usr/bin/env python
"""Loop bodies of a kernel in the built .so: instruction count and mnemonic histogram (offline check before GPU time)."""
import re, subprocess, sys
from collections import Counter
so, pat = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = raw.split("Function : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name: continue
    ins = []
    for l in b.splitlines():
        m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    print(name[:80], len(ins), "instructions")
    seen = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r'BRA.*0x([0-9a-f]+)', t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                j = addr[tgt]
                n = i - j + 1
                if n < minlen or n > 1500: continue
                c = Counter(re.sub(r'^@!?U?P\d+\s+', '', x).split()[0].split('.')[0] for _, x in ins[j:i + 1])
                print(f"  loop {ins[j][0]:#x}-{ins[i][0]:#x}: {n} instr;", dict(c.most_common(16)))
