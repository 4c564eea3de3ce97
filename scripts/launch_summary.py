#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel
count, total and mean device time, and (with --steady N) the share of each kernel
over the last N launches of the probe kernel's steps.  Usage:
    python scripts/launch_summary.py gpurun_out/<tag>/launches.csv [--tail-from KERNEL_SUBSTR]"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    else:
        raise SystemExit("no CSV header in " + path)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1000.0, "us": v, "ms": v * 1000.0, "s": v * 1e6}.get(r[ui], v)
        seq.append((name, v))
    return seq


def table(seq, title):
    agg = collections.OrderedDict()
    for n, v in seq:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"## {title}: {len(seq)} launches, {tot:.1f} us device time")
    print("| launches | total us | us/launch | share | kernel |")
    print("|---:|---:|---:|---:|---|")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {100 * a[1] / tot:.1f}% | `{n[:100]}` |")
    print()


def main():
    seq = load(sys.argv[1])
    table(seq, "all launches")
    if "--tail-from" in sys.argv:
        key = sys.argv[sys.argv.index("--tail-from") + 1]
        idx = [i for i, (n, _) in enumerate(seq) if key in n]
        if idx:
            # steady state: from the last but one occurrence of the step's first kernel to the end
            nth = int(sys.argv[sys.argv.index("--nth") + 1]) if "--nth" in sys.argv else 2
            table(seq[idx[-nth]:], f"last {nth} steps (from `{key}`)")


if __name__ == "__main__":
    main()
