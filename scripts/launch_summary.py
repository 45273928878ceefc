#!/usr/bin/env python
"""Per-kernel shares from an ncu launch list (`--metrics gpu__time_duration.sum --csv`): usage
launch_summary.py launches.csv STEPS ["title"].  Times under ncu are serialised and cold-cache: read the shares."""
import collections
import csv
import sys


def main(path, steps, title=""):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    t, n = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        t[r[ik]] += float(r[iv].replace(",", ""))
        n[r[ik]] += 1
    tot = sum(t.values())
    print(f"# {title or path}: {sum(n.values())} launches over {steps} steps, {tot / steps / 1e3:.1f} us/step (serialised)")
    for k, v in t.most_common():
        print(f"{v / steps / 1e3:9.1f} us/step {100 * v / tot:5.1f}% {n[k] / steps:6.1f} launches/step  {k[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else "")
