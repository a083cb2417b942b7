"""Print selected raw metrics of the first kernel in an .ncu-rep (development aid): python tools/ncu_raw.py X.ncu-rep [regex]"""
import csv
import re
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, v = r[0], r[2] if len(r) > 2 else r[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"gpu__time_duration.sum|smsp__inst_executed.sum$|inst_executed.avg.per_cycle_elapsed|registers_per_thread|dram__bytes_(read|write).sum$|"
                 r"pipe_(alu|fma|fmaheavy|fmalite|lsu|xu|uniform)\.sum.pct_of_peak_sustained_active|issue_stalled.*per_issue_active|"
                 r"lts__t_sector_hit_rate.pct|sm__warps_active.avg.pct|gpu__dram_throughput.avg.pct|launch__grid_size|sm__throughput.avg.pct|"
                 r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|smsp__issue_active.avg.pct")
for k, x in zip(h, v):
    if pat.search(k):
        print("%-90s %s" % (k, x))
