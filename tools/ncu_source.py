#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export (SASS view) into windows of
instructions with their dominant stall reasons.
usage: python tools/ncu_source.py src.csv [window]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
win = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
for w0 in range(0, len(data), win):
    chunk = data[w0:w0 + win]
    ns = sum(int(r[ix["# Samples"]] or 0) for r in chunk)
    st = {c: sum(int(r[ix[c]] or 0) for r in chunk) for c in stall_cols}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    ops = {}
    for r in chunk:
        op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
        if op.startswith("@"):
            op = r[ix["Source"]].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    topops = ",".join(f"{k}{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
    print(f"[{w0:5d}-{w0+len(chunk):5d}) {100*ns/tot:5.1f}%  " +
          " ".join(f"{k[6:]}={100*v/max(ns,1):.0f}%" for k, v in top) + "   " + topops)
