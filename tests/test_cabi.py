"""CPU: the C-ABI library loads, exports every symbol include/polargpu.h declares, and refuses to run without a GPU
(no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "polargpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from polardecoding_b200 import load_library
    lib = load_library()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libpolargpu.so lacks %s" % s


def test_presets_match_reference_programs():
    from polardecoding_b200.capi import preset, PROGRAMS
    from oracle_lib import Oracle
    for prog in PROGRAMS:
        p = preset(prog)
        o = Oracle(prog)
        assert (p.N, p.K, p.crc_bits) == (o.N, o.K, o.r)
        if prog.startswith(("SCL", "CASCL")):
            assert p.list_size == o.L == 8
        if prog.startswith("BP"):
            assert p.iter_max == o.iters
        assert p.crc_poly == o.code.crc_poly and p.crc_systematic == o.code.crc_systematic
    with pytest.raises(ValueError):
        preset("SC_bitRev_buggy")


def test_no_cpu_fallback(have_gpu):
    from polardecoding_b200 import Engine
    from polardecoding_b200.capi import preset
    if have_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device|failed"):
        Engine("SC_128")
    # argument errors are reported before any device is touched
    p = preset("SC_128")
    p.N = 100
    with pytest.raises(RuntimeError, match="power of two"):
        Engine(params=p)


def test_product_does_not_reference_the_oracle():
    """the product path must not import, link or execute anything under oracle/"""
    pkg = os.path.join(ROOT, "polardecoding_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "polar_oracle" not in txt and "oracle/" not in txt and "oracle_lib" not in txt, os.path.join(dp, f)


def test_host_logic_partition_and_merge():
    from polardecoding_b200 import load_library, PgCounters
    lib = load_library()
    lib.pg_partition.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    s, c = C.c_uint64(), C.c_uint64()
    got = []
    for r in range(4):
        assert lib.pg_partition(1000, 256, 4, r, 600, C.byref(s), C.byref(c)) == 0
        got.append((s.value, c.value))
    assert got == [(1000, 256), (1256, 256), (1512, 88), (1768, 0)]
    assert lib.pg_partition(0, 0, 4, 0, 1, C.byref(s), C.byref(c)) != 0


def test_fp32_table_increments_land_on_the_literals():
    """polar_common.cuh forms the 8-level table in fp32 as a running sum of increments (FMA pipe); each partial sum must
    equal the fp32 literal the reference's table holds (0.05f ... 0.65f), so the form is bit-identical to compare/select."""
    import re
    import numpy as np
    src = open(os.path.join(ROOT, "polardecoding_b200", "csrc", "polar_common.cuh")).read()
    body = src[src.index("__device__ __forceinline__ float tbl8<float>(float a)"):src.index("tbl8_select_f32")]
    incs = [np.uint32(int(h, 16)).view(np.float32) for h in re.findall(r"__int_as_float\((0x[0-9a-f]+)\)", body)]
    thr = [float(t) for t in re.findall(r"NB, ([0-9.]+)f \*", body)]
    assert thr == [4.5, 2.252, 1.508, 1.05, 0.71, 0.433, 0.196] and len(incs) == 7
    lits = [np.float32(v) for v in (0.05, 0.15, 0.25, 0.35, 0.45, 0.55, 0.65)]
    acc = np.float32(0)
    for inc, lit in zip(incs, lits):
        acc = np.float32(acc + inc)          # fmaf(1.0f, inc, acc) rounds once, like this add
        assert acc == lit, (acc, lit)
