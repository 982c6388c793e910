#!/usr/bin/env python
"""Digest of an `ncu --page source --csv` export: total samples per stall reason, the hottest instructions,
and samples grouped by opcode.  usage: ncu_src.py file.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
tot = collections.Counter(); byop = collections.Counter(); insts = collections.Counter()
samples = 0
for r in body:
    s = int(r[ix["# Samples"]] or 0); samples += s
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].strip() else "?"
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    byop[op] += s; insts[op] += int(r[ix["Instructions Executed"]] or 0)
    for h in stalls: tot[h] += int(r[ix[h]] or 0)
print("samples", samples, "instructions executed", sum(insts.values()))
print("stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / samples) for k, v in tot.most_common(10)))
print("by opcode (samples%, inst%):")
ti = sum(insts.values())
for op, s in byop.most_common(22): print("  %-22s %5.1f%%  %5.1f%%" % (op, 100.0 * s / samples, 100.0 * insts[op] / ti))
print("hottest instructions:")
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]; s = int(r[ix["# Samples"]])
    why = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print("  %5d %5.2f%%  %-70s %s" % (i, 100.0 * s / samples, r[ix["Source"]].strip()[:70], " ".join("%s=%d" % (n, c) for c, n in why)))
