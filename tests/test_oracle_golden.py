"""CPU: the oracle (oracle/polar_oracle.c) against golden vectors produced by the COMPILED REFERENCE
(tools/make_golden.py) and against the reference's own captured result files (KATs K1, K2, K4, K5 of
SURVEY.md section 4).  This is what pins the oracle; the GPU tests then compare the kernels with it."""
import json
import os

import numpy as np
import pytest

from oracle_lib import Oracle

def norm(t):
    return " ".join(t.split())


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PROGS = ["SC_128", "SC_1024", "SC_128_fag", "SCL_128", "SCL_128_fag", "CASCL_128", "SCL_1024", "CASCL_1024_L8",
         "CASCL_1024_sys", "BP_128", "BP_128_fag", "BP_1024"]


def load(prog):
    z = np.load(os.path.join(GOLD, prog + ".npz"))
    N = int(z["N"])
    u = np.unpackbits(z["u"], axis=1, bitorder="little")[:, :N]
    uh = np.unpackbits(z["u_hat"], axis=1, bitorder="little")[:, :N]
    return z, z["llr"].astype(np.float64), u, uh


@pytest.mark.parametrize("prog", PROGS)
def test_oracle_matches_reference_vectors(prog):
    z, llr, u, uh = load(prog)
    o = Oracle(prog)
    assert o.N == int(z["N"]) and o.K == int(z["K"]) and o.nI == int(z["nI"]) and o.L == int(z["L"]) and o.iters == int(z["iters"])
    assert (o.I == z["I"]).all()
    got, _ = o.decode(llr)
    assert (got == uh).all()


def test_oracle_bpr_statistic():
    z, llr, u, uh = load("BPr_128")
    o = Oracle("BPr_128")
    got, E = o.bpr(llr, u, z["samples"])
    assert (got == uh).all()
    assert (E == z["E"]).all()


def test_kat_captured_results():
    kat = json.load(open(os.path.join(GOLD, "kat.json")))
    # K1: SC_128 as shipped; every line of the author's capture (run AND error-bit columns)
    k = kat["K1_SC_128"]
    res = Oracle("SC_128").simulate_ref(0, k["ebn0"], k["target"], k["seed"])
    assert [r[0] for r in res] == k["run"] == [252, 364, 707, 1505, 4766, 15386, 53195]
    cap = norm(kat["captures"]["myResult_128/SC128out.txt"])  # the author's capture (older build: no error-bit line)
    out = norm(k["stdout"])                                    # the reference binary re-run in the build container
    for (run, eb, ebit), snr in zip(res, k["ebn0"]):
        assert "bSNR = %.2f error block = %d run = %d" % (snr, eb, run) in cap
        assert "bSNR = %.2f error block = %d run = %d BLER = %f Error bit = %d BER = %f" % (snr, eb, run, eb / run, ebit, ebit / 64 / run) in out
    # K2: SCL_128 L=8
    k = kat["K2_SCL_128"]
    res = Oracle("SCL_128").simulate_ref(1, k["ebn0"], k["target"], k["seed"])
    assert [r[0] for r in res] == k["run"] == [204, 411, 845, 1953]
    cap = norm(kat["captures"]["myResult_128/SCL128out_errblock50.dat"])
    for (run, eb, _), snr in zip(res, k["ebn0"]):
        assert "L = 8 bSNR = %.2f error block = %d run = %d" % (snr, eb, run) in cap


def scl128_capture_runs(kat):
    """{L: [(bSNR, run)...]} from the author's capture myResult_128/SCL128out_errblock50.dat (SCL_128.c with SEED = 1024, 50 errors,
    one run of the program per list size, `#define L` edited by hand: SCL_128.c:16)"""
    import re
    runs = {}
    for m in re.finditer(r"L = (\d+)\s+bSNR = ([0-9.]+)\s+error block = 50\s+run = (\d+)", kat["captures"]["myResult_128/SCL128out_errblock50.dat"]):
        runs.setdefault(int(m.group(1)), []).append((float(m.group(2)), int(m.group(3))))
    return runs


def test_kat_scl128_every_list_size():
    """K2 for all five list sizes of the capture (L = 2, 4, 8, 16, 32), frame-exact `run` columns up to 3.0 dB"""
    kat = json.load(open(os.path.join(GOLD, "kat.json")))
    runs = scl128_capture_runs(kat)
    assert sorted(runs) == [2, 4, 8, 16, 32] and all(len(v) == 6 for v in runs.values())
    o = Oracle("SCL_128")
    for L, pts in runs.items():
        pts = pts[:5]
        res = o.simulate_ref(1, [p[0] for p in pts], 50, 1024, L=L)
        assert [r[0] for r in res] == [p[1] for p in pts], (L, res)


def test_kat_cascl():
    kat = json.load(open(os.path.join(GOLD, "kat.json")))
    # K4: CASCL_128, the SEED = 8392 block of myResult_128/CASCL_128_L8.txt (first three points: seconds on one core)
    k = kat["K4_CASCL_128"]
    res = Oracle("CASCL_128").simulate_ref(2, k["ebn0"][:3], k["target"], k["seed"])
    assert [r[0] for r in res] == k["run"][:3] == [843, 1712, 4782]
    cap = norm(kat["captures"]["myResult_128/CASCL_128_L8.txt"])
    assert "SEED = 8392" in cap
    for (run, eb, _), snr in zip(res, k["ebn0"]):
        assert "bSNR = %.2f error block = %d run = %d" % (snr, eb, run) in cap
    # K5: CASCL_1024_L8, SEED = 1242 block of myResult_1024/CASCL_L8.dat: run = 246, 1381 at 1.0 / 1.5 dB, 100 errors
    res = Oracle("CASCL_1024_L8").simulate_ref(2, [1.0, 1.5], 100, 1242)
    assert [r[0] for r in res] == [246, 1381]


def test_primitives_and_frames():
    o = Oracle("CASCL_1024_L8")
    lib = o.lib
    # table boundaries of CHK / PHI (SC_128.c:293-307): value at a threshold belongs to the upper interval
    for t, lo, hi in [(0.196, 0.65, 0.55), (0.433, 0.55, 0.45), (0.71, 0.45, 0.35), (1.05, 0.35, 0.25), (1.508, 0.25, 0.15), (2.252, 0.15, 0.05), (4.5, 0.05, 0.0)]:
        assert lib.po_phi(np.nextafter(t, 0), 0) == lo and lib.po_phi(t, 0) == hi
        assert lib.po_phi(-t, 0) == hi + t and lib.po_phi(t, 1) == hi + t
    assert lib.po_chk(0.0, 0.0) == 0.0 and lib.po_chk(999.0, 1.25) == 1.25 and lib.po_chk(-999.0, 1.25) == -1.25
    assert lib.po_chk(1.0, 1.0) == 1.0 + (0.15 - 0.65)
    # CRC-24 frames produced by po_make_u are multiples of g(D) and the systematic variant carries the payload
    import ctypes as C
    pn = np.zeros(63, dtype=np.int32)
    lib.po_pn63(pn.ctypes.data_as(C.POINTER(C.c_int)))
    assert "".join(map(str, pn)) == "100000100001100010100111101000111001001011011101100110101011111"
    for prog in ("CASCL_1024_L8", "CASCL_1024_sys", "CASCL_128"):
        oo = Oracle(prog)
        u, _ = oo.frames_ref_stream(3.0, 3, seed=1)
        for f in range(3):
            cw = np.ascontiguousarray(u[f, oo.I], dtype=np.int32)
            assert lib.po_crc_check(C.byref(oo.code), cw.ctypes.data_as(C.POINTER(C.c_int))) == 1
            cw[5] ^= 1
            assert lib.po_crc_check(C.byref(oo.code), cw.ctypes.data_as(C.POINTER(C.c_int))) == 0
    oo = Oracle("CASCL_1024_sys")
    u, _ = oo.frames_ref_stream(3.0, 1, seed=1)
    assert (u[0, oo.I[24:24 + 63]] == pn).all()
