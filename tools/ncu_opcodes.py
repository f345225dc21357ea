"""Opcode histogram (warp instructions executed, stall samples) of one kernel from `ncu -i X.ncu-rep --page source --csv`."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
isrc, iex, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
c, s = Counter(), Counter()
tot = 0
for r in rows[hi + 1:]:
    if len(r) <= iex or not r[iex].isdigit():
        continue
    ex, st = int(r[iex]), int(r[ist] or 0)
    parts = r[isrc].split()
    op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
    c[op] += ex; s[op] += st; tot += ex
print("total warp instructions", tot)
for op, v in c.most_common(28):
    print(f"{op:10s} {v / tot * 100:6.2f}%  {v:>12d}  stall samples {s[op]}")
