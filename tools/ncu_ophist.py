"""Opcode histogram (weighted by executed warp-instructions) + stall samples from an ncu source-page csv:
   ncu -i X.ncu-rep --page source --csv > X.src.csv ; python tools/ncu_ophist.py X.src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci, ce, cs = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
hist, stall, tot, stot = collections.Counter(), collections.Counter(), 0, 0
reasons = collections.Counter()
for r in rows[2:]:
    try:
        n = int(float(r[ce]))
    except (ValueError, IndexError):
        continue
    toks = r[ci].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    hist[op] += n
    tot += n
    s = int(float(r[cs] or 0))
    stall[op] += s
    stot += s
    for i, h in stall_cols:
        try:
            reasons[h] += int(float(r[i] or 0))
        except ValueError:
            pass
print("total warp-instructions", tot, " stall samples", stot)
for op, n in hist.most_common(26):
    print(f"{op:10s} {n / tot * 100:6.2f}% of instr   {stall[op] / max(stot, 1) * 100:6.2f}% of stall samples")
print("stall reasons:", ", ".join(f"{k[6:]}={v / max(stot, 1) * 100:.1f}%" for k, v in reasons.most_common(8)))
