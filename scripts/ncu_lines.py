#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname = ""; hdr = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0].isdigit() and r[2] == "-":   # a CUDA source line aggregate
        ie = hdr.index("Instructions Executed"); ns = hdr.index("# Samples")
        d = {h: r[i] for i, h in enumerate(hdr)}
        out.append((int(r[ie]), int(r[ns]), fname, int(r[0]), r[1].strip()[:100], d))
tot = sum(x[0] for x in out); tots = sum(x[1] for x in out)
print("total warp-instructions", tot, "samples", tots)
stall_keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for x in sorted(out, key=lambda x: -x[1])[:top]:
    st = sorted(((int(x[5][k] or 0), k[6:]) for k in stall_keys), reverse=True)[:3]
    print(f"{100*x[0]/tot:5.1f}%i {100*x[1]/tots:5.1f}%s {x[2]}:{x[3]:<4d} {x[4]:<100s} {' '.join(f'{k}={v}' for v,k in st if v)}")
if len(sys.argv) > 3:
    # ranges "a-b,c-d" over track.cu lines: share of instructions / samples
    for rg in sys.argv[3].split(","):
        a, b = map(int, rg.split("-"))
        i = sum(x[0] for x in out if x[2].startswith("track.cu") and a <= x[3] <= b)
        s_ = sum(x[1] for x in out if x[2].startswith("track.cu") and a <= x[3] <= b)
        print(f"track.cu {a}-{b}: {100*i/tot:5.1f}% inst {100*s_/tots:5.1f}% samples")
    i = sum(x[0] for x in out if not x[2].startswith("track.cu")); s_ = sum(x[1] for x in out if not x[2].startswith("track.cu"))
    print(f"other files: {100*i/tot:5.1f}% inst {100*s_/tots:5.1f}% samples")
