"""Per-source-line share of executed instructions and stall samples from `ncu -i X.ncu-rep --page source --csv --print-source=cuda`
(development aid; usage: python tools/ncu_lines.py file.csv [top])."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
hdr, out, cur = None, [], ""
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        try:
            ln, ie, smp = int(r[0]), int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        out.append((cur, ln, ie, smp, r[1][:120]))
tot, ts = sum(o[2] for o in out), sum(o[3] for o in out)
print("total warp instructions", tot, "samples", ts)
for f, ln, ie, smp, s in sorted(out, key=lambda x: -x[2])[:top]:
    print("%-18s %4d inst %6.2f%%  samples %5.2f%%  %s" % (f, ln, 100 * ie / tot, 100 * smp / max(ts, 1), s))
