"""Turn the round's ncu captures (gpurun_out/*.ncu-rep, launches csv) into the tracked text summaries under profiles/.
  python tools/summarize_profiles.py TAG        (TAG names the round/version, e.g. r1b)"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__time_duration\.sum|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|launch__(block_size|grid_size|"
                  r"occupancy_limit_(blocks|registers|shared_mem)|registers_per_thread)|lts__t_sector_hit_rate\.pct|sm__cycles_elapsed\.avg|"
                  r"sm__inst_executed\.avg\.per_cycle_elapsed|sm__inst_executed_pipe_(alu|fma|fmaheavy|fp16|lsu|xu)\.sum\.pct_of_peak_sustained_active|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__inst_executed\.sum|smsp__thread_inst_executed_per_inst_executed\.ratio|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active)$")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    return dict(zip(r[0], r[2] if len(r) > 2 else r[1])), dict(zip(r[0], r[1]))


try:  # frames per launch of the profiled command = frames per step of the bench line of the same round
    _b = json.load(open(os.path.join(ROOT, "gpurun_out", "bench_%s.json" % tag)))
    FRAMES = {"cascl": _b["config"]["frames_per_step"], "bp": _b["bp_1024"]["frames_per_step"], "bph2": _b["bp_1024"]["frames_per_step"],
              "bp64": _b.get("f64_parity_mode", {}).get("bp_1024", {}).get("frames_per_step")}
except Exception:
    FRAMES = {}
traffic = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the kernel inside `python bench.py --steps 2 --warmup 3 --no-cpu`, "
                    "ncu --set full --clock-control none (profiles/%s_*_summary.txt)" % tag}
for key, title in (("cascl", "CA-SCL N=1024 L=8 fp32 list kernel"), ("cascl64", "CA-SCL N=1024 L=8 fp64 (bit-exact) list kernel, one quarter-grid launch"),
                   ("bp", "BP N=1024 fp32, 100 sweeps"), ("bp64", "BP N=1024 fp64 (bit-exact), 100 sweeps"), ("bph2", "BP N=1024 packed-half mode (optional flag)")):
    rep = os.path.join(ROOT, "gpurun_out", "prof_%s_%s.ncu-rep" % (key, tag))
    if not os.path.exists(rep):
        rep = os.path.join(ROOT, "gpurun_out", "prof_%s_r1_final2.ncu-rep" % key)
    if not os.path.exists(rep):
        continue
    v, units = raw(rep)
    lines = ["%s -- one launch inside `python bench.py --steps 2 --warmup 3 --no-cpu`" % title,
             "source: ncu --set full --clock-control none --import-source on; file %s" % os.path.basename(rep), "",
             "%-90s %s" % ("Kernel Name", v.get("Kernel Name", ""))]
    for k in sorted(v):
        if KEEP.match(k) and v[k] not in ("", "0"):
            lines.append("%-90s %s %s" % (k, v[k], units.get(k, "")))
    open(os.path.join(ROOT, "profiles", "%s_%s_summary.txt" % (tag, key)), "w").write("\n".join(lines) + "\n")

    def gb(x, u):
        return float(x) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
    tot = gb(v["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) + gb(v["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
    traffic[key] = {"kernel": v.get("Kernel Name", ""), "dram_bytes_per_launch": tot, "grid": int(v["launch__grid_size"]),
                    "time_ms_under_ncu": float(v["gpu__time_duration.sum"]) * {"ms": 1.0, "us": 1e-3, "s": 1e3}.get(units["gpu__time_duration.sum"], 1.0),
                    "frames_per_launch": FRAMES.get(key) if key != "cascl64" else int(v["launch__grid_size"]) * 4}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "%s_traffic.json" % tag), "w"), indent=1)

# launch list
lc = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
if os.path.exists(lc):
    rows = [r for r in csv.reader(l for l in open(lc) if not l.startswith("==")) if r]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        t = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    out = ["launch list of `python bench.py --steps 2 --warmup 3 --no-cpu` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("%-70s n=%4d total %10.3f ms share %5.1f%% avg %.3f ms" % (k[:70], n, t, 100 * t / tot, t / n))
    open(os.path.join(ROOT, "profiles", "%s_launches_summary.txt" % tag), "w").write("\n".join(out) + "\n")
    import shutil
    shutil.copy(lc, os.path.join(ROOT, "profiles", "%s_launches.csv" % tag))
print(open(os.path.join(ROOT, "profiles", "%s_traffic.json" % tag)).read())
