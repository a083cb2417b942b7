"""ctypes access to the CPU warp emulator build of the CUDA kernels (tests/emu).  Test infrastructure only: it lets the
`-m "not gpu"` suite check kernel logic (schedules, pointer bookkeeping, collectives, memory layouts) against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
FLAVORS = {"": "", "chunk": "-DPOLAR_VIRT=0 -DPOLAR_CHUNK=8 -DEMU_FEW"}   # build flavour -> extra compile options (see tests/emu/Makefile)

_libs = {}


def lib(flavor=""):
    if flavor not in _libs:
        cmd = ["make", "-C", EMU_DIR, "-j4"]
        if flavor:
            cmd += ["FLAVOR=" + flavor, "EMU_EXTRA=" + FLAVORS[flavor]]
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
        l = C.CDLL(os.path.join(EMU_DIR, "_build" + ("_" + flavor if flavor else ""), "libpolar_emu.so"))
        vp = C.c_void_p
        l.emu_list_decode.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.c_ulonglong, vp, vp, vp, C.c_int, C.c_int, C.c_int,
                                      C.c_uint, vp, vp, vp, C.c_int]
        _libs[flavor] = l
    return _libs[flavor]


def crc_masks(I, N, r, poly):
    """u-domain syndrome masks exactly as csrc/api.cu builds them (CRcheck, CASCL_1024_L8.c:569-598)."""
    W = N // 32
    nI = len(I)
    low = poly & ((1 << r) - 1)
    cur, rem = 1, []
    for _ in range(nI):
        rem.append(cur)
        cur <<= 1
        if (cur >> r) & 1:
            cur = (cur & ((1 << r) - 1)) ^ low
    m = np.zeros((max(r, 1), W), dtype=np.uint32)
    for i in range(nI):
        for b in range(r):
            if (rem[i] >> b) & 1:
                m[b, I[i] >> 5] |= np.uint32(1 << (I[i] & 31))
    return m


def list_decode(oracle, llr, L, use_crc, f64=True, grid=1, coop=True, count_from=0, tm=1, flavor=""):
    """Decode llr (B,N) with the emulated list kernel; oracle supplies the code (I, inI, r, crc_poly).
    tm: 0 = layout without tensor memory, 1 = the dispatcher's choice, 2 / 3 = a tensor-memory layout forced for any type (stages 3..5 / stage 6 only; N >= 256).
    -> (u_hat (B,N) int32, frame_info (B,) uint32, collectives executed)"""
    N = oracle.N
    n = int(np.log2(N))
    W = N // 32
    llr = np.ascontiguousarray(llr, dtype=np.float64 if f64 else np.float32).reshape(-1, N)
    B = llr.shape[0]
    info = np.zeros(W, dtype=np.uint32)
    cnt = np.zeros(W, dtype=np.uint32)
    for i, p in enumerate(oracle.I):
        info[p >> 5] |= np.uint32(1 << (p & 31))
        if i >= count_from:
            cnt[p >> 5] |= np.uint32(1 << (p & 31))
    masks = crc_masks([int(x) for x in oracle.I], N, oracle.r, int(oracle.code.crc_poly))
    first = int(np.argmax(oracle.inI)) if oracle.inI.any() else N
    out = np.zeros((B, W), dtype=np.uint32)
    fi = np.zeros(B, dtype=np.uint32)
    coll = C.c_ulonglong(0)
    rc = lib(flavor).emu_list_decode(n, L, int(f64), llr.ctypes.data, B, info.ctypes.data, cnt.ctypes.data, masks.ctypes.data, oracle.r,
                               int(use_crc), (first // 4) if coop else 0, grid, out.ctypes.data, fi.ctypes.data, C.byref(coll), tm)
    assert rc == 0, "configuration (n=%d, L=%d, tm=%d) is not compiled into the emulator (rc %d)" % (n, L, tm, rc)
    bits = ((out[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(B, N).astype(np.int32)
    return bits, fi, coll.value
