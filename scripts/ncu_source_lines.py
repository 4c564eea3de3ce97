#!/usr/bin/env python
"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` by source line:
executed warp instructions and stall samples per line (first captured launch only).
Usage: ncu_source_lines.py src.csv [top_n]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg, stall, text = collections.Counter(), collections.Counter(), {}
cur, hdr, seen = None, None, set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        if cur in seen and cur == first:
            break            # second launch starts
        if not seen:
            first = cur
        seen.add(cur)
        hdr = None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        i_ie, i_st = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or r[2] != "-" or not r[0].isdigit():
        continue
    ie = int(r[i_ie]) if r[i_ie].isdigit() else 0
    st = int(r[i_st]) if r[i_st].isdigit() else 0
    key = (cur, int(r[0]))
    agg[key] += ie
    stall[key] += st
    text[key] = r[1][:100]
tot, tot_st = sum(agg.values()), max(1, sum(stall.values()))
print("total warp instructions", tot, "stall samples", tot_st)
byfile = collections.Counter()
for (f, ln), ie in agg.items():
    byfile[f] += ie
print({f: round(v / tot * 100, 1) for f, v in byfile.items()})
for (f, ln), ie in sorted(agg.items(), key=lambda x: -(x[1] / tot + stall[x[0]] / tot_st))[:top]:
    print(f"{f}:{ln:4d} instr {ie / tot * 100:5.1f}%  stall {stall[(f, ln)] / tot_st * 100:5.1f}%  {text[(f, ln)]}")
