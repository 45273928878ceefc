#!/usr/bin/env python
"""Per-kernel SASS mnemonic digest of the shipped libcomo_b200.so (cuobjdump -sass), for profiles/: which kernels
use the TMA unit (UBLKCP / UTMALDG), mbarriers (SYNCS), cp.async (LDGSTS), the FP64 tensor path (DMMA), packed fp32
(FFMA2) -- and that none of the tcgen05 family (UTC*MMA, LDTM/STTM: no f64 kind exists) is expected here."""
import re, subprocess, sys
from collections import Counter, OrderedDict
so = sys.argv[1] if len(sys.argv) > 1 else "como_b200/libcomo_b200.so"
WATCH = ["DMMA", "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "FFMA2",
         "FMUL2", "FADD2", "DFMA", "FFMA", "ATOMS", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "BAR", "CCTL"]
raw = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
out = OrderedDict()
for blk in raw.split("Function : ")[1:]:
    name = blk.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    c, n = Counter(), 0
    for l in blk.splitlines():
        m = re.search(r'/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', l)
        if m:
            n += 1
            c[m.group(1)] += 1
    out[dem] = (n, c)
print(f"# SASS digest of {so}: instructions per kernel, watched mnemonics (exact opcode stem)")
print(f"{'kernel':48s} {'instr':>6s}  " + " ".join(f"{w:>7s}" for w in WATCH if any(o[1].get(w) for o in out.values())))
cols = [w for w in WATCH if any(o[1].get(w) for o in out.values())]
for k, (n, c) in out.items():
    print(f"{k[-48:]:48s} {n:6d}  " + " ".join(f"{c.get(w, 0):7d}" for w in cols))
tot = Counter()
for n, c in out.values():
    tot.update(c)
print("\n# totals: " + ", ".join(f"{w}={tot.get(w, 0)}" for w in WATCH))
