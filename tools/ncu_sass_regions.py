"""SASS in address order with its source line, executed count and stall samples, from
`ncu -i X.ncu-rep --page source --csv --print-source=cuda,sass` (development aid).
  python tools/ncu_sass_regions.py file.csv [dump]   -> per-region summary (regions split at branch targets / big gaps), or a full dump"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; cur_file = ""; cur_line = 0; out = []
for r in rows:
    if r and r[0] == "File Path": cur_file = r[1].split("/")[-1]
    if r and r[0] == "Line No": hdr = r; continue
    if not hdr or len(r) != len(hdr): continue
    if r[0] != "":
        try: cur_line = int(r[0])
        except ValueError: pass
        continue
    if r[2].startswith("0x"):
        ie = int(r[hdr.index("Instructions Executed")]); smp = int(r[hdr.index("# Samples")])
        out.append((int(r[2], 16), cur_file, cur_line, r[3].strip(), ie, smp))
out.sort()
tot_i = sum(o[4] for o in out); tot_s = sum(o[5] for o in out)
if len(sys.argv) > 2:
    for a, f, l, s, ie, smp in out:
        print("%x %-16s %4d %-60s %6.3f%% %6.3f%%" % (a - out[0][0], f[:16], l, s[:60], 100.0 * ie / tot_i, 100.0 * smp / tot_s))
else:
    # regions = maximal runs with the same executed count bucket
    reg = []; start = 0
    for i in range(1, len(out) + 1):
        if i == len(out) or abs(out[i][4] - out[i - 1][4]) > 0.02 * max(out[i][4], out[i - 1][4], 1):
            seg = out[start:i]
            reg.append((seg[0][0] - out[0][0], len(seg), sum(s[4] for s in seg), sum(s[5] for s in seg), seg[0][4],
                        sorted(set(s[2] for s in seg if s[1].startswith("list_decode")))[:6]))
            start = i
    for off, n, ie, smp, per, lines in reg:
        if ie * 200 > tot_i or smp * 200 > tot_s:
            print("@%6x n=%4d exec/inst %10d inst %6.2f%% samples %6.2f%%  lines %s" % (off, n, per, 100.0 * ie / tot_i, 100.0 * smp / tot_s, lines))
