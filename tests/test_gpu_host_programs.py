"""GPU: the drop-in C host programs (polardecoding_b200/host/bin/<prog>).  With --rng ref they build frames with the
reference's own generator on the host and decode on the GPU; their stdout must then be byte-identical to what the
unmodified reference program prints (KAT captures in tests/golden/kat.json, taken from the compiled reference)."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "polardecoding_b200", "host", "bin")
KAT = json.load(open(os.path.join(ROOT, "tests", "golden", "kat.json")))


def run(prog, *args, stdin=None):
    r = subprocess.run([os.path.join(BIN, prog)] + list(args), stdin=stdin or subprocess.DEVNULL, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return r.stdout


def test_sc128_reproduces_reference_stdout_exactly():
    out = run("SC_128", "--rng", "ref")
    assert out == KAT["K1_SC_128"]["stdout"]


def test_scl128_reproduces_reference_stdout_exactly():
    out = run("SCL_128", "--rng", "ref")
    assert out == KAT["K2_SCL_128"]["stdout"]


def test_cascl128_reproduces_reference_stdout():
    k = KAT["K4_CASCL_128"]
    out = run("CASCL_128", "--rng", "ref", "--seed", "8392", "--ebn0", "1.0:0.5:2.0")
    want = "".join(l + "\n" for l in k["stdout"].splitlines()[:4])          # SEED line + 3 points
    assert out == want


def test_cascl1024_kat_k5():
    out = run("CASCL_1024_L8", "--rng", "ref", "--seed", "1242", "--ble", "100")
    assert out.splitlines()[0] == "SEED = 1242"
    assert "L = 8\tbSNR = 1.00\terror block = 100\trun = 246\tBLER = 4065.040650e-4" in out      # myResult_1024/CASCL_L8.dat
    assert "L = 8\tbSNR = 1.50\terror block = 100\trun = 1381\tBLER = 724.112962e-4" in out


def test_stdin_matrix_is_accepted_and_philox_mode_runs():
    fn = "\n".join(" ".join("1" if (i & j) == j else "0" for j in range(128)) for i in range(128)) + "\n"
    p = subprocess.run([os.path.join(BIN, "BP_128"), "--ebn0", "2.0", "--ble", "20", "--seed", "3", "--verbose"], input=fn, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    lines = p.stdout.splitlines()
    assert lines[0] == "SEED = 3" and lines[1].startswith("bSNR = 2.00\terror block = 20\trun = ")
    assert "Mframes/s" in p.stderr
    out = run("SC_1024", "--ebn0", "2.0", "--ble", "10", "--real", "f64")
    assert out.startswith("bSNR = 2.00\terror block = 10\trun = ")


def test_bpr128_reproduces_reference_stdout():
    """BPr_128 with time() forced to 945: SEED line and the first two Eb/N0 points, byte for byte (run, six E rows, BLER/BER)"""
    out = run("BPr_128", "--rng", "ref", "--seed", "945", "--ebn0", "1.0:0.5:1.5")
    assert out == KAT["K_BPr_128"]["stdout"]
    out = run("BPr_128", "--ebn0", "2.0", "--ble", "50", "--seed", "3")
    assert "After 80 iterations:" in out and "K * BER" in out


# ---- round 2: every list size of the author's captures, K3, and the print formats of the remaining programs -----------------
FIRST = json.load(open(os.path.join(ROOT, "tests", "golden", "kat_firstlines.json")))


def capture_blocks(text):
    """{L: [line, ...]} of a capture that holds one block of result lines per list size"""
    import re
    blocks = {}
    for line in text.replace("\r", "").split("\n"):
        m = re.match(r"L = (\d+)\tbSNR = ", line)
        if m:
            blocks.setdefault(int(m.group(1)), []).append(line)
    return blocks


@pytest.mark.parametrize("L", [2, 4, 16, 32])
def test_scl128_capture_every_list_size(L):
    """myResult_128/SCL128out_errblock50.dat: SCL_128.c with `#define L` edited per run (SCL_128.c:16), SEED = 1024, six points"""
    want = capture_blocks(KAT["captures"]["myResult_128/SCL128out_errblock50.dat"])[L]
    assert len(want) == 6
    out = run("SCL_128", "--rng", "ref", "--L", str(L), "--ebn0", "1.0:0.5:3.5")
    assert out == "".join(l + "\n" for l in want)


@pytest.mark.parametrize("L,points", [(8, 5), (2, 4), (4, 4), (16, 4), (32, 3)])
def test_scl1024_capture_k3(L, points):
    """K3 = myResult_1024/SCL1024out.dat (SEED = 1024): run = 227, 1026, 5867, 21575, 178842 for L = 8 -- all five points -- and the
    first points of the other four list sizes (L = 16 / 32 at N = 1024: the 64-bit pointer word, one or two frames per warp)"""
    want = capture_blocks(FIRST["capture_SCL1024out"])[L][:points]
    if L == 8:
        assert [int(l.split("run = ")[1].split()[0]) for l in want] == [227, 1026, 5867, 21575, 178842]
    out = run("SCL_1024", "--rng", "ref", "--L", str(L), "--ebn0", "1.0:0.5:%.1f" % (1.0 + 0.5 * (points - 1)))
    assert out == "".join(l + "\n" for l in want)


@pytest.mark.parametrize("prog", ["SC_1024", "SC_128_fag", "SCL_1024", "SCL_128_fag", "BP_128", "BP_128_fag", "BP_1024", "CASCL_1024_sys"])
def test_first_lines_equal_the_compiled_reference(prog):
    """print formats of the programs not covered above (e.g. BP_1024.c:255-257, CASCL_1024_sys.c:832-835, SCL_128_fag.c:256-259): the
    first stdout lines of the compiled reference (tools/make_kat_firstlines.py) against the drop-in program with the same seed"""
    import re
    k = FIRST[prog]
    want = k["stdout"]
    snrs = [float(x) for x in re.findall(r"bSNR = ([0-9.]+)", want)]
    assert snrs, "golden has no result line"
    args = ["--rng", "ref", "--seed", str(k["seed"]), "--ebn0", "%.1f:0.5:%.1f" % (snrs[0], snrs[-1])]
    if "ble" in k:
        args += ["--ble", str(k["ble"])]
    out = run(prog, *args)
    n = len(want.splitlines())
    assert "".join(l + "\n" for l in out.splitlines()[:n]) == want
