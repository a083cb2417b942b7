"""CPU: tools/results.py parses the author's captured result files (stored verbatim in tests/golden/kat.json) and the
reference binaries' stdout, and compares result sets with a confidence interval."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import results  # noqa: E402

KAT = json.load(open(os.path.join(ROOT, "tests", "golden", "kat.json")))


def test_parse_captures_and_reference_stdout():
    cap = results.parse(KAT["captures"]["myResult_128/CASCL_128_L8.txt"])           # spaces, several SEED blocks
    blk = [r for r in cap if r["seed"] == 8392]
    assert [r["run"] for r in blk] == [843, 1712, 4782, 21054, 105937] and all(r["err"] == 200 and r["L"] == 8 for r in blk)
    sc = results.parse(KAT["captures"]["myResult_128/SC128out.txt"])                 # CRLF + blank lines, no L column
    assert [r["run"] for r in sc] == [252, 364, 707, 1505, 4766, 15386, 53195] and sc[0]["L"] is None
    scl = results.parse(KAT["captures"]["myResult_128/SCL128out_errblock50.dat"])    # L = 2..32 blocks
    assert sorted({r["L"] for r in scl}) == [2, 4, 8, 16, 32]
    ours = results.parse(KAT["K4_CASCL_128"]["stdout"])                               # what the binaries print (tabs)
    assert [r["run"] for r in ours] == [843, 1712, 4782, 21054, 105937] and ours[0]["seed"] == 8392


def test_compare_uses_pooled_reference_and_flags_outliers():
    ref = results.parse(KAT["captures"]["myResult_128/CASCL_128_L8.txt"])
    same = results.compare(ref, results.parse(KAT["K4_CASCL_128"]["stdout"]))
    assert len(same) == 5 and all(c["inside"] for c in same)
    bad = results.compare(ref, results.parse("L = 8\tbSNR = 2.00\terror block = 200\trun = 1000\tBLER = x\n"))
    assert len(bad) == 1 and not bad[0]["inside"] and bad[0]["z"] > 5
