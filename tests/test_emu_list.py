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


TM_CASES = [  # program or (N, K), L, use_crc, frames, Eb/N0 -- every one with N >= 256 (a tensor-memory layout needs a scratch stage above it)
    ("CASCL_1024_L8", 8, 1, 9, 1.0),      # 9 frames: three warps of a CTA busy, the fourth idle, a ragged last group
    ("SC_1024", 1, 0, 40, 1.5),           # L = 1: 32 frames per warp, no pointer words
    ("CASCL_1024_L8", 32, 1, 3, 1.0),     # 64-bit pointer words
    ((256, 100), 4, 0, 20, 1.0),          # smallest N: the prefix subtree is raised to the first scratch stage
    ((1024, 1000), 8, 0, 4, 4.0),         # first information bit among the first leaves: no cooperative prefix at all
    ((1024, 40), 8, 0, 4, 1.0),           # prefix capped at N/8 groups, resumes outside the subtree
]


@pytest.mark.parametrize("prog,L,crc,B,snr", TM_CASES)
def test_emulated_tensor_memory_layout_equals_oracle_f64(prog, L, crc, B, snr):
    """The layouts with LLR stages in tensor memory (four-warp CTAs, own-row loads + shuffle for cloned pointers, the
    copied-out prefix blocks, the raised prefix subtree), forced for fp64 in the emulator so that the oracle checks them bit for bit."""
    o = Oracle(prog) if isinstance(prog, str) else Oracle(N=prog[0], K=prog[1], L=L)
    rng = np.random.default_rng(77 + 7 * L + B)
    llr = awgn_llr(rng, o.N, B, snr)
    want, aux = o.decode(llr, kind=("sc" if L == 1 else ("cascl" if crc else "scl")), L=L)
    for grid, tm in ((1, 2), (2, 3)):   # tm 2: stages 3..5 in tensor memory (the product layout); 3: stage 6 only, 3..5 in smem
        got, fi, coll = emu_lib.list_decode(o, llr, L, crc, f64=True, grid=grid, tm=tm)
        assert (got == want * o.inI[None, :]).all(), "%d of %d frames differ (grid %d, tm %d)" % (int((got != want * o.inI[None, :]).any(1).sum()), B, grid, tm)
        if L > 1:
            assert (((fi >> 16) & 3) == aux).all()


@pytest.mark.parametrize("prog,L,crc", [("CASCL_1024_L8", 8, 1), ("SCL_1024", 2, 0), ("SC_1024", 1, 0)])
def test_emulated_fp32_tensor_memory_kernel_equals_plain_fp32_kernel(prog, L, crc):
    """fp32 with stages 3..5 in tensor memory against the fp32 product layout (no tensor memory): same arithmetic, so the
    decisions and flags must be identical word for word, also on frames where fp32 differs from fp64."""
    o = Oracle(prog)
    B = 11 if L > 1 else 70
    llr = awgn_llr(np.random.default_rng(5 + L), o.N, B, 1.0, dtype=np.float32)
    a, fa, _ = emu_lib.list_decode(o, llr, L, crc, f64=False, grid=2, tm=0)
    b, fb, _ = emu_lib.list_decode(o, llr, L, crc, f64=False, grid=1, tm=2)
    assert (a == b).all() and (fa == fb).all()


@pytest.mark.parametrize("prog,L,crc,B,snr", [("CASCL_1024_L8", 8, 1, 6, 1.0), ((1024, 1000), 8, 0, 4, 4.0), ((512, 300), 8, 0, 6, 2.0), ("SC_1024", 1, 0, 33, 1.5)])
def test_emulated_chunked_chain_equals_oracle_f64(prog, L, crc, B, snr):
    """the optional chunked chain (-DPOLAR_CHUNK=8: top g-layers produced eight rows per half at a time, each chunk consumed at once by
    the f-layer below; off in the product) in a second emulator build: same decisions as the oracle"""
    o = Oracle(prog) if isinstance(prog, str) else Oracle(N=prog[0], K=prog[1], L=L)
    llr = awgn_llr(np.random.default_rng(17 + B), o.N, B, snr)
    want, aux = o.decode(llr, kind=("sc" if L == 1 else ("cascl" if crc else "scl")), L=L)
    got, fi, _ = emu_lib.list_decode(o, llr, L, crc, f64=True, grid=2, tm=0, flavor="chunk")
    assert (got == want * o.inI[None, :]).all()
    if L > 1:
        assert (((fi >> 16) & 3) == aux).all()
