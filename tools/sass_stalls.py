#!/usr/bin/env python
"""Top stall sites of an ncu report's SASS source page (one section per captured kernel).
usage: sass_stalls.py report.ncu-rep|source_page.csv [topN]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
if rep.endswith(".csv"):   # an already exported source page (tools/r02_*.sh export it on the box: reports are ~16 MB each)
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for sec in sections:
    hdr, data = sec["hdr"], sec["data"]
    iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[iS]) for r in data)
    print('=' * 30, sec["name"][:110]); print('total samples', tot)
    agg = {}
    for r in data:
        for i in stall_cols:
            agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i])
    print('stall totals', sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    top = sorted(enumerate(data), key=lambda t: -int(t[1][iS]))[:topn]
    for idx, r in sorted(top):
        st = {hdr[i][6:]: int(r[i]) for i in stall_cols if int(r[i]) > 0}
        st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(idx, r[iS].rjust(6), r[iEx].rjust(8), r[iSrc].strip()[:64].ljust(64), st)
