"""Regenerates include/polar_q_table.h (the 3GPP TS 38.212 Table 5.3.1.2-1 reliability sequence)
from the table the reference programs hard-code, and checks that every reference program uses the
same sequence (the N=128 programs: that table filtered to indices < 128).  Needs /root/reference;
the generated header is committed, so this only has to run when the header is to be re-derived."""
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def q_of(path):
    s = open(path, "rb").read().decode("latin1").replace("\r", "")
    m = re.search(r"Q\[N\]\s*=\s*\{([^}]*)\}", s)
    return [int(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()]


q1024 = q_of(REF + "/SC_1024.c")
q128 = q_of(REF + "/SC_128.c")
assert sorted(q1024) == list(range(1024)) and [x for x in q1024 if x < 128] == q128
for f in ["CASCL_1024_L8", "SCL_1024", "BP_1024", "CASCL_1024_sys", "BPRGA_1024", "BPRGA_1024_W"]:
    assert q_of("%s/%s.c" % (REF, f)) == q1024, f
for f in ["SCL_128", "CASCL_128", "BP_128", "BPr_128", "SC_128_fag", "SCL_128_fag", "BP_128_fag", "BPDEGA_128",
          "BPRGA_128", "BPRGA_128_W", "BPRGA_128_M", "BPRGA_128_allbit"]:
    assert q_of("%s/%s.c" % (REF, f)) == q128, f
rows = ["    " + ", ".join("%4d" % v for v in q1024[i:i + 16]) + "," for i in range(0, 1024, 16)]
hdr = open("include/polar_q_table.h").read()
head = hdr[: hdr.index("= {\n") + 4]
open("include/polar_q_table.h", "w").write(head + "\n".join(rows) + "\n};\n#endif\n")
print("include/polar_q_table.h regenerated")
