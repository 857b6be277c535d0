#!/usr/bin/env python
"""Top stall sites of an ncu source-page CSV (SASS view): address, samples, executed count, instruction."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = sum(int(r[iS] or 0) for r in data)
totI = sum(int(r[iN] or 0) for r in data)
print("total samples", tot, "total warp-instructions", totI)
top = sorted(range(len(data)), key=lambda i: -int(data[i][iS] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[iS]):7d} {100*int(r[iS])/tot:5.1f}% exec={r[iN]:>10s} {r[iSrc].strip()[:70]:70s} {st}")
