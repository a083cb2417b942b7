"""Kernel LOGIC of csrc/list_decode.cu on the CPU: the CUDA source is compiled unchanged for a warp emulator (tests/emu)
and its fp64 decisions must equal the oracle's bit for bit -- the same bar the `-m gpu` parity tests apply on the device.
(Reference functions: SCdecode SC_128.c:395-460, SCLdecode SCL_1024.c:547-680, CASCL CASCL_1024_L8.c:601-761.)"""
import numpy as np
import pytest

import emu_lib
from oracle_lib import Oracle, awgn_llr

CASES = [
    # program, N/K override (None = preset), L, use_crc, frames, Eb/N0
    ("SC_128", 1, 0, 40, 2.0),
    ("SC_1024", 1, 0, 33, 1.5),
    ("SCL_128", 2, 0, 18, 1.0),
    ("SCL_128", 8, 0, 9, 1.0),
    ("SCL_128", 32, 0, 3, 1.0),
    ("CASCL_128", 8, 1, 12, 1.0),
    ("SCL_1024", 8, 0, 5, 1.0),
    ("CASCL_1024_L8", 8, 1, 9, 1.0),
    ("CASCL_1024_L8", 2, 1, 17, 1.5),
    ("CASCL_1024_L8", 16, 1, 3, 1.0),
    ("CASCL_1024_L8", 32, 1, 2, 1.0),
    ("CASCL_1024_sys", 8, 1, 5, 1.5),
]


@pytest.mark.parametrize("prog,L,crc,B,snr", CASES)
def test_emulated_kernel_equals_oracle_f64(prog, L, crc, B, snr):
    o = Oracle(prog)
    rng = np.random.default_rng(1000 + 7 * L + B)
    llr = awgn_llr(rng, o.N, B, snr)
    want, aux = o.decode(llr, kind=("sc" if L == 1 else ("cascl" if crc else "scl")), L=L)
    got, fi, coll = emu_lib.list_decode(o, llr, L, crc, f64=True, grid=2)
    assert coll > 0 or L == 1
    assert (got == want * o.inI[None, :]).all(), "%d of %d frames differ" % (int((got != want * o.inI[None, :]).any(1).sum()), B)
    if L > 1:
        assert (((fi >> 16) & 3) == aux).all()  # tie / CRC-fail flags as the oracle reports them


def test_emulated_kernel_without_the_cooperative_prefix():
    o = Oracle("CASCL_1024_L8")
    llr = awgn_llr(np.random.default_rng(3), o.N, 5, 1.0)
    want, _ = o.decode(llr)
    got, _, _ = emu_lib.list_decode(o, llr, 8, 1, f64=True, coop=False)
    assert (got == want * o.inI[None, :]).all()


@pytest.mark.parametrize("N,K,L", [(32, 16, 4), (64, 20, 8), (256, 100, 4), (512, 300, 8), (512, 500, 16), (1024, 1000, 8), (1024, 40, 8)])
def test_emulated_kernel_other_code_shapes(N, K, L):
    """rates and lengths the reference programs do not use: short and long frozen prefixes, every storage class of a stage"""
    o = Oracle(N=N, K=K, L=L)
    B = 6 if N < 1024 else 4
    llr = awgn_llr(np.random.default_rng(N + K), N, B, 1.0 if K < N // 2 else 4.0)
    want, _ = o.decode(llr, kind="scl", L=L)
    got, _, _ = emu_lib.list_decode(o, llr, L, 0, f64=True)
    assert (got == want * o.inI[None, :]).all()


def test_emulated_fp32_kernel_rarely_differs_from_fp64():
    o = Oracle("CASCL_1024_L8")
    llr = awgn_llr(np.random.default_rng(11), o.N, 12, 2.0, dtype=np.float32)
    want, _ = o.decode(llr)
    got, _, _ = emu_lib.list_decode(o, llr, 8, 1, f64=False)
    assert int((got != want * o.inI[None, :]).any(1).sum()) == 0
