"""Stall-reason samples per SASS address range from `ncu --page source --csv --print-source=cuda,sass` (development aid).
  python tools/ncu_stalls_by_region.py file.csv lo:hi[:name] ...   (hex offsets from the first instruction)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; out = []
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] == "" and r[2].startswith("0x"):
        out.append((int(r[2], 16), r))
out.sort()
base = out[0][0]
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[hdr.index("# Samples")]) for _, r in out)
regs = sys.argv[2:] or ["0:fffff:all"]
for spec in regs:
    p = spec.split(":"); lo, hi = int(p[0], 16), int(p[1], 16); name = p[2] if len(p) > 2 else spec
    acc = {c: 0 for c in cols}; ie = 0; n = 0
    for a, r in out:
        if lo <= a - base < hi:
            n += 1; ie += int(r[hdr.index("Instructions Executed")])
            for c in cols: acc[c] += int(r[c])
    s = sum(acc.values())
    print("%-10s n=%4d samples %5.2f%% :" % (name, n, 100.0 * s / tot), " ".join("%s %.1f%%" % (hdr[c][6:], 100.0 * v / max(s, 1)) for c, v in sorted(acc.items(), key=lambda x: -x[1])[:7]))
