"""Executed warp instructions by SASS opcode from `ncu -i X.ncu-rep --page source --csv --print-source=cuda,sass` (development aid)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, ops = None, collections.Counter()
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] == "" and r[2].startswith("0x"):
        ie = int(r[hdr.index("Instructions Executed")])
        t = r[3].split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.rstrip(";")] += ie
tot = sum(ops.values())
print("total", tot)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print("%-24s %6.2f%%" % (op, 100 * n / tot))
