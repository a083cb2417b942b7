"""CPU: the CRC parity table in the reference's file format (/root/reference/CRC_6.dat) through pg_crc_table_load.
The fixture is not a copy of the reference's file: tools/make_crc_dat.py regenerates it from g(D) = D^6 + D^5 + 1 and the result
is byte-identical to the reference's file (SHA-256 recorded here from /root/reference/CRC_6.dat; compared directly when the
reference tree is present)."""
import ctypes as C
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF_SHA256 = "7065f2fe8cc21a177409812d103d504a61b249d35e866fe466b6638288965128"


def load(path, K, r):
    from polardecoding_b200 import load_library
    lib = load_library()
    lib.pg_crc_table_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.c_void_p]
    poly = C.c_uint64(0)
    rows = np.zeros(K, dtype=np.uint32)
    rc = lib.pg_crc_table_load(path.encode(), K, r, C.byref(poly), rows.ctypes.data)
    return rc, poly.value, rows


def test_generated_file_is_the_reference_file():
    from make_crc_dat import crc_table_bytes
    data = b"\xff\xfe" + crc_table_bytes(64, 6, 0x61)
    assert hashlib.sha256(data).hexdigest() == REF_SHA256
    ref = os.path.join(os.environ.get("POLAR_REF", "/root/reference"), "CRC_6.dat")
    if os.path.exists(ref):
        assert open(ref, "rb").read() == data


def test_loader_reads_utf16_and_ascii_and_recovers_the_polynomial(tmp_path):
    from make_crc_dat import crc_table_bytes
    p16 = tmp_path / "CRC_6.dat"
    p16.write_bytes(b"\xff\xfe" + crc_table_bytes(64, 6, 0x61))
    rc, poly, rows = load(str(p16), 64, 6)
    assert rc == 0 and poly == 0x61                      # PG_CRC6_POLY: D^6 + D^5 + 1 (CASCL_128.c:18)
    assert rows[0] == 0b100001 and rows[1] == 0b100011   # D^6 mod g = D^5 + 1; D^7 mod g = D^5 + D + 1
    # rows are what the engine derives from the polynomial for its systematic encoder: D^(r+i) mod g
    cur, want = 1, []
    for i in range(6 + 64):
        want.append(cur)
        cur <<= 1
        if cur >> 6:
            cur = (cur & 63) ^ 0b100001
    assert (rows == np.array(want[6:], dtype=np.uint32)).all()
    # the same table as plain ASCII, and the CRC-24 table of CASCL_1024_sys.c (512 x 24) in the same format
    pa = tmp_path / "ascii.dat"
    pa.write_text(crc_table_bytes(64, 6, 0x61).decode("utf-16-le"))
    assert load(str(pa), 64, 6)[:2] == (0, 0x61)
    p24 = tmp_path / "crc24.dat"
    p24.write_bytes(b"\xff\xfe" + crc_table_bytes(512, 24, 0x1B2B117))
    assert load(str(p24), 512, 24)[:2] == (0, 0x1B2B117)


def test_loader_rejects_what_is_not_a_crc_table(tmp_path):
    from make_crc_dat import crc_table_bytes
    good = crc_table_bytes(64, 6, 0x61).decode("utf-16-le")
    bad = tmp_path / "bad.dat"
    bad.write_text(good[:40] + ("0" if good[40] == "1" else "1") + good[41:])   # one flipped coefficient
    assert load(str(bad), 64, 6)[0] != 0
    bad.write_text(good)
    assert load(str(bad), 63, 6)[0] != 0                                         # wrong shape
    bad.write_text(good.replace("1", "2", 1))
    assert load(str(bad), 64, 6)[0] != 0                                         # not 0/1
    assert load(str(tmp_path / "missing.dat"), 64, 6)[0] != 0
