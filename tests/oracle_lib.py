"""ctypes access to the test oracle (oracle/_build/libpolar_oracle.so, the CPU restatement) and,
when it has been built in a container that holds /root/reference, to the compiled reference
harnesses (oracle/_ref/libref_<prog>.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "_build", "libpolar_oracle.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

PO_MAX_N = 2048


class PoCode(C.Structure):
    _fields_ = [("N", C.c_int), ("n", C.c_int), ("K", C.c_int), ("r", C.c_int), ("nI", C.c_int),
                ("crc_systematic", C.c_int), ("crc_poly", C.c_uint64),
                ("I", C.c_int * PO_MAX_N), ("inI", C.c_uint8 * PO_MAX_N)]


class PoRng(C.Structure):
    _fields_ = [("v", C.c_uint64)]


class PoPoint(C.Structure):
    _fields_ = [("run", C.c_long), ("err_block", C.c_long), ("err_bit", C.c_long)]


_port = None


def build_port():
    if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "polar_oracle.c")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
    return PORT_SO


def port():
    global _port
    if _port is None:
        lib = C.CDLL(build_port())
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        lib.po_code_preset.argtypes = [C.POINTER(PoCode), C.c_char_p, ip, ip]
        lib.po_code_init.argtypes = [C.POINTER(PoCode), C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int]
        lib.po_chk.restype = C.c_double
        lib.po_chk.argtypes = [C.c_double, C.c_double]
        lib.po_phi.restype = C.c_double
        lib.po_phi.argtypes = [C.c_double, C.c_int]
        lib.po_sc_decode.argtypes = [C.POINTER(PoCode), dp, ip]
        lib.po_scl_decode.argtypes = [C.POINTER(PoCode), C.c_int, C.c_int, dp, ip, ip]
        lib.po_bp_decode.argtypes = [C.POINTER(PoCode), C.c_int, dp, ip, ip]
        lib.po_bpr_decode.argtypes = [C.POINTER(PoCode), C.c_int, dp, ip, ip, ip, C.c_int, ip]
        lib.po_rng_seed.argtypes = [C.POINTER(PoRng), C.c_uint64]
        lib.po_rng_uniform.restype = C.c_double
        lib.po_rng_uniform.argtypes = [C.POINTER(PoRng)]
        lib.po_rng_normal_pair.argtypes = [C.POINTER(PoRng), C.c_double, dp, dp]
        lib.po_pn63.argtypes = [ip]
        lib.po_make_u.argtypes = [C.POINTER(PoCode), ip, C.c_int, ip]
        lib.po_polar_encode.argtypes = [C.POINTER(PoCode), ip, ip]
        lib.po_crc_check.argtypes = [C.POINTER(PoCode), ip]
        lib.po_simulate_ref.argtypes = [C.POINTER(PoCode), C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                        C.c_int, C.POINTER(PoRng), ip, C.POINTER(PoPoint)]
        _port = lib
    return _port


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Oracle:
    """The restated decoder for one of the reference's named programs."""

    def __init__(self, prog=None, N=None, K=None, r=0, crc_poly=0, crc_systematic=0, L=1, iters=0):
        self.lib = port()
        self.code = PoCode()
        if prog is not None:
            l_, it_ = C.c_int(), C.c_int()
            rc = self.lib.po_code_preset(C.byref(self.code), prog.encode(), C.byref(l_), C.byref(it_))
            assert rc == 0, prog
            self.L, self.iters = l_.value, it_.value
        else:
            rc = self.lib.po_code_init(C.byref(self.code), N, K, r, crc_poly, crc_systematic)
            assert rc == 0
            self.L, self.iters = L, iters
        self.prog = prog
        self.N, self.K, self.r, self.nI = self.code.N, self.code.K, self.code.r, self.code.nI
        self.I = np.array(self.code.I[: self.nI], dtype=np.int32)
        self.inI = np.array(self.code.inI[: self.N], dtype=np.uint8)

    def kind(self):
        p = self.prog or ""
        if p.startswith("SC_"):
            return "sc"
        if p.startswith("SCL_"):
            return "scl"
        if p.startswith("CASCL_"):
            return "cascl"
        if p.startswith("BPr"):
            return "bpr"
        return "bp"

    def decode(self, llr, kind=None, L=None, iters=None):
        """llr: (B,N) float64 -> (u_hat (B,N) int32, aux (B,) int32: SCL flags / BP fixed-point sweep)."""
        kind = kind or self.kind()
        llr = np.ascontiguousarray(llr, dtype=np.float64).reshape(-1, self.N)
        B = llr.shape[0]
        out = np.zeros((B, self.N), dtype=np.int32)
        aux = np.zeros(B, dtype=np.int32)
        L = L or self.L
        iters = iters or self.iters
        one = C.c_int()
        for f in range(B):
            if kind == "sc":
                self.lib.po_sc_decode(C.byref(self.code), _dp(llr[f]), _ip(out[f]))
            elif kind in ("scl", "cascl"):
                self.lib.po_scl_decode(C.byref(self.code), L, int(kind == "cascl"), _dp(llr[f]), _ip(out[f]), C.byref(one))
                aux[f] = one.value
            else:
                self.lib.po_bp_decode(C.byref(self.code), iters, _dp(llr[f]), _ip(out[f]), C.byref(one))
                aux[f] = one.value
        return out, aux

    def bpr(self, llr, u_true, samples, iters=None):
        llr = np.ascontiguousarray(llr, dtype=np.float64).reshape(-1, self.N)
        u_true = np.ascontiguousarray(u_true, dtype=np.int32).reshape(-1, self.N)
        B = llr.shape[0]
        samples = np.ascontiguousarray(samples, dtype=np.int32)
        E = np.zeros((len(samples), self.code.n + 1), dtype=np.int32)
        out = np.zeros((B, self.N), dtype=np.int32)
        for f in range(B):
            self.lib.po_bpr_decode(C.byref(self.code), iters or self.iters, _dp(llr[f]), _ip(u_true[f]), _ip(out[f]),
                                   _ip(samples), len(samples), _ip(E))
        return out, E

    def simulate_ref(self, decoder, ebn0_list, target, seed, count_from=0, L=None, iters=None):
        """The reference main() loop with the reference's generator: list of (run, err_block, err_bit)."""
        g = PoRng()
        self.lib.po_rng_seed(C.byref(g), seed)
        m = C.c_int(0)
        res = []
        for e in ebn0_list:
            pt = PoPoint()
            self.lib.po_simulate_ref(C.byref(self.code), decoder, L or self.L, iters or self.iters, float(e), target,
                                     count_from, C.byref(g), C.byref(m), C.byref(pt))
            res.append((pt.run, pt.err_block, pt.err_bit))
        return res

    def frames_ref_stream(self, ebn0_db, nframes, seed, m0=0):
        """u (B,N), llr (B,N) produced exactly as the reference main() would from `seed` (fresh generator)."""
        g = PoRng()
        self.lib.po_rng_seed(C.byref(g), seed)
        pn = np.zeros(63, dtype=np.int32)
        self.lib.po_pn63(_ip(pn))
        sigma = 10 ** (ebn0_db / -20.0)
        u = np.zeros((nframes, self.N), dtype=np.int32)
        x = np.zeros(self.N, dtype=np.int32)
        llr = np.zeros((nframes, self.N), dtype=np.float64)
        a, b = C.c_double(), C.c_double()
        m = m0
        for f in range(nframes):
            self.lib.po_make_u(C.byref(self.code), _ip(pn), m, _ip(u[f]))
            self.lib.po_polar_encode(C.byref(self.code), _ip(u[f]), _ip(x))
            y = np.zeros(self.N)
            for i in range(0, self.N, 2):
                self.lib.po_rng_normal_pair(C.byref(g), sigma, C.byref(a), C.byref(b))
                y[i] = (1 if x[i] == 0 else -1) + a.value
                y[i + 1] = (1 if x[i + 1] == 0 else -1) + b.value
            llr[f] = 2 * y / sigma / sigma
            m = (m + self.K % 63) % 63
        return u, llr


def have_ref(prog):
    return os.path.exists(os.path.join(REF_DIR, "libref_%s.so" % prog))


class RefHarness:
    """The reference's own decoder (compiled from the unmodified source) on shared LLRs."""

    def __init__(self, prog):
        self.lib = C.CDLL(os.path.join(REF_DIR, "libref_%s.so" % prog))
        self.lib.ref_chk.restype = C.c_double
        self.lib.ref_chk.argtypes = [C.c_double, C.c_double]
        self.lib.ref_decode_llr.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.lib.ref_normal_pair.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        self.lib.ref_rng_restart.argtypes = [C.c_uint64]
        self.lib.ref_init()
        self.N = self.lib.ref_param_N()
        self.K = self.lib.ref_param_K()
        self.nI = self.lib.ref_param_nI()
        self.L = self.lib.ref_param_L()
        self.iters = self.lib.ref_param_iter()
        self.I = np.zeros(self.nI, dtype=np.int32)
        self.inI = np.zeros(self.N, dtype=np.int32)
        self.lib.ref_info_set(_ip(self.I), _ip(self.inI))
        if hasattr(self.lib, "ref_phi"):
            self.lib.ref_phi.restype = C.c_double
            self.lib.ref_phi.argtypes = [C.c_double, C.c_int]

    def decode(self, llr, truth=None):
        llr = np.ascontiguousarray(llr, dtype=np.float64).reshape(-1, self.N)
        out = np.zeros(llr.shape, dtype=np.int32)
        tp = None
        if truth is not None:
            truth = np.ascontiguousarray(truth, dtype=np.int32)
            tp = _ip(truth)
        self.lib.ref_decode_llr(_dp(llr), llr.shape[0], _ip(out), tp)
        return out


def awgn_llr(rng, N, B, ebn0_db, x=None, dtype=np.float64):
    """BPSK/AWGN LLRs the way the reference forms them (sigma^2 = 1/(Eb/N0), rate 1/2 baked in)."""
    sigma = 10 ** (ebn0_db / -20.0)
    s = 1.0 - 2.0 * (x if x is not None else np.zeros((B, N)))
    y = s + sigma * rng.standard_normal((B, N))
    return (2 * y / sigma / sigma).astype(dtype).astype(np.float64)
