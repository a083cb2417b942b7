"""Result-file tooling: parse result lines as the reference programs and the drop-in host programs
print them -- including the author's captured files inside myResult_*.zip (UTF-16 or UTF-8, CRLF, tabs or spaces) -- and
compare two result sets point by point with a 95 % interval (both are fixed-error-count Monte-Carlo estimates).
Replaces the hand-pasted numbers of the reference's SCL_1024.py / plot_SCL.py.

  python tools/results.py parse  FILE|ZIP:member ...
  python tools/results.py compare REF GPU        (each FILE or ZIP:member)
"""
import math
import re
import sys
import zipfile

LINE = re.compile(r"(?:L\s*=\s*(?P<L>\d+)\s+)?bSNR\s*=\s*(?P<snr>[\d.]+)\s+(?:error block\s*=\s*(?P<err>\d+)\s+)?run\s*=\s*(?P<run>\d+)")
SEED = re.compile(r"SEED\s*=\s*(\d+)(?:\s+error block\s*=\s*(\d+))?")


def read_text(spec):
    if ":" in spec and spec.split(":", 1)[0].lower().endswith(".zip"):
        z, member = spec.split(":", 1)
        raw = zipfile.ZipFile(z).read(member)
    else:
        raw = open(spec, "rb").read()
    if raw[:2] in (b"\xff\xfe", b"\xfe\xff"):
        return raw.decode("utf-16")
    return raw.decode("utf-8", errors="replace")


def parse(text):
    """-> list of dicts {seed, L, snr, err, run, bler}; `err` falls back to the header's 'error block = N' (CASCL_1024_sys)."""
    rows, seed, hdr_err = [], None, None
    for line in text.replace("\r", "").split("\n"):
        m = SEED.search(line)
        if m:
            seed = int(m.group(1))
            hdr_err = int(m.group(2)) if m.group(2) else None
        m = LINE.search(line)
        if not m:
            continue
        err = int(m.group("err")) if m.group("err") else hdr_err
        run = int(m.group("run"))
        rows.append({"seed": seed, "L": int(m.group("L")) if m.group("L") else None, "snr": float(m.group("snr")), "err": err, "run": run,
                     "bler": (err / run) if err is not None and run else None})
    return rows


def compare(ref_rows, gpu_rows):
    out = []
    for g in gpu_rows:
        cands = [r for r in ref_rows if abs(r["snr"] - g["snr"]) < 1e-9 and (g["L"] is None or r["L"] in (None, g["L"])) and r["bler"]]
        if not cands or not g["bler"]:
            continue
        # pool the reference's seeds for this point
        e = sum(r["err"] for r in cands)
        n = sum(r["run"] for r in cands)
        rb = e / n
        z = (g["bler"] - rb) / (rb * math.sqrt(1.0 / e + 1.0 / g["err"]))
        out.append({"snr": g["snr"], "L": g["L"], "ref_bler": rb, "ref_err": e, "gpu_bler": g["bler"], "gpu_err": g["err"], "z": z, "inside": abs(z) < 1.96})
    return out


def main(argv):
    if len(argv) >= 2 and argv[0] == "parse":
        for spec in argv[1:]:
            for r in parse(read_text(spec)):
                print("%s\tseed=%s\tL=%s\tEb/N0=%.2f\terr=%s\trun=%d\tBLER=%s" % (spec, r["seed"], r["L"], r["snr"], r["err"], r["run"], "%.4g" % r["bler"] if r["bler"] else "-"))
        return 0
    if len(argv) == 3 and argv[0] == "compare":
        res = compare(parse(read_text(argv[1])), parse(read_text(argv[2])))
        for c in res:
            print("Eb/N0 %.2f L=%s: reference %.4g (%d errors) vs %.4g (%d errors): z = %+.2f %s" % (c["snr"], c["L"], c["ref_bler"], c["ref_err"], c["gpu_bler"], c["gpu_err"], c["z"], "ok" if c["inside"] else "OUTSIDE 95%"))
        return 0 if all(c["inside"] for c in res) else 1
    print(__doc__)
    return 2


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
