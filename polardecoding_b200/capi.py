"""ctypes binding of include/polargpu.h.  Thin: every call goes straight to libpolargpu.so.
Raises if the library is missing -- there is deliberately no fallback path."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POLARGPU_LIB") or os.path.join(HERE, "libpolargpu.so")  # override: development A/B builds only

PROGRAMS = ["SC_128", "SC_1024", "SC_128_fag", "SCL_128", "SCL_1024", "SCL_128_fag", "CASCL_128", "CASCL_1024_L8",
            "CASCL_1024_sys", "BP_128", "BP_1024", "BP_128_fag", "BPr_128"]
EXTRA_PROGRAMS = ["CASCL_128_sys"]   # a variant the reference documents by its data (CRC_6.dat) and result files only

PG_DEC_SC, PG_DEC_SCL, PG_DEC_CASCL, PG_DEC_BP = 0, 1, 2, 3
PG_REAL_F64, PG_REAL_F32, PG_REAL_H2 = 0, 1, 2
PG_DATA_PN63, PG_DATA_PHILOX = 0, 1


class PgParams(C.Structure):
    _fields_ = [("N", C.c_int), ("K", C.c_int), ("crc_bits", C.c_int), ("crc_poly", C.c_uint64),
                ("crc_systematic", C.c_int), ("decoder", C.c_int), ("list_size", C.c_int), ("iter_max", C.c_int),
                ("bp_early_stop", C.c_int), ("real", C.c_int), ("data_mode", C.c_int), ("count_from", C.c_int),
                ("device", C.c_int), ("seed", C.c_uint64), ("rank", C.c_int), ("nranks", C.c_int), ("llr_clip", C.c_float)]


class PgCounters(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("err_blocks", C.c_uint64), ("err_bits", C.c_uint64),
                ("tie_frames", C.c_uint64), ("crc_fail", C.c_uint64), ("bp_sweeps", C.c_uint64),
                ("reserved", C.c_uint64 * 2)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k in ("frames", "err_blocks", "err_bits", "tie_frames", "crc_fail", "bp_sweeps")}


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C polardecoding_b200/csrc); there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.pg_version.restype = C.c_char_p
    lib.pg_params_preset.argtypes = [C.POINTER(PgParams), C.c_char_p]
    lib.pg_create.argtypes = [C.POINTER(PgParams), C.POINTER(vp)]
    lib.pg_destroy.argtypes = [vp]
    lib.pg_destroy.restype = None
    lib.pg_last_error.argtypes = [vp]
    lib.pg_last_error.restype = C.c_char_p
    lib.pg_info_set.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_uint8)]
    lib.pg_decode_llr.argtypes = [vp, vp, C.c_int, C.c_size_t, vp, vp]
    lib.pg_decode_llr_packed.argtypes = [vp, vp, C.c_int, C.c_size_t, vp, vp]
    lib.pg_decode_llr_counted.argtypes = [vp, vp, C.c_int, C.c_size_t, vp, vp, C.POINTER(PgCounters), vp]
    lib.pg_decode_llr_device.argtypes = [vp, vp, C.c_int, C.c_size_t, vp, vp]
    lib.pg_decode_count_device.argtypes = [vp, vp, C.c_int, C.c_size_t, vp, vp, vp]
    lib.pg_counters_read.argtypes = [vp, C.POINTER(PgCounters), C.c_int]
    lib.pg_channel_device.argtypes = [vp, C.c_double, C.c_uint64, C.c_size_t, vp, vp]
    lib.pg_channel.argtypes = [vp, C.c_double, C.c_uint64, C.c_size_t, vp, vp]
    lib.pg_simulate.argtypes = [vp, C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(PgCounters)]
    lib.pg_simulate_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.pg_simulate_batch.argtypes = [vp, C.c_double, C.c_uint64, C.c_size_t, C.POINTER(PgCounters), vp]
    lib.pg_bpr_config.argtypes = [vp, C.POINTER(C.c_int), C.c_int]
    lib.pg_bpr_read.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.pg_bpr_reset.argtypes = [vp]
    lib.pg_comm_unique_id.argtypes = [vp]
    lib.pg_comm_init.argtypes = [vp, vp]
    lib.pg_allreduce_counters.argtypes = [vp, C.POINTER(PgCounters)]
    lib.pg_wave_frames.argtypes = [vp]
    lib.pg_wave_frames.restype = C.c_uint64
    lib.pg_sync.argtypes = [vp]
    lib.pg_stream.argtypes = [vp]
    lib.pg_stream.restype = vp
    lib.pg_kernel_launches.argtypes = [vp]
    lib.pg_kernel_launches.restype = C.c_uint64
    lib.pg_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    _lib = lib
    return lib


def preset(program):
    p = PgParams()
    rc = load_library().pg_params_preset(C.byref(p), program.encode())
    if rc != 0:
        raise ValueError("unknown program %r" % program)
    return p


class Engine:
    """One pg_ctx.  `program` picks the reference program whose defaults are used; keyword arguments
    override fields of pg_params (real='f64'|'f32', list_size, iter_max, bp_early_stop, seed, data_mode, ...)."""

    def __init__(self, program=None, params=None, **over):
        self.lib = load_library()
        p = params if params is not None else preset(program)
        for k, v in over.items():
            if k == "real":
                v = {"f64": PG_REAL_F64, "f32": PG_REAL_F32, "h2": PG_REAL_H2}.get(v, v)
            setattr(p, k, v)
        self.params = p
        self.ctx = C.c_void_p()
        rc = self.lib.pg_create(C.byref(p), C.byref(self.ctx))
        if rc != 0:
            raise RuntimeError("pg_create failed (%d): %s" % (rc, self.lib.pg_last_error(None).decode()))
        self.N, self.K, self.nI = p.N, p.K, p.K + p.crc_bits
        self.f64 = p.real == PG_REAL_F64
        I = (C.c_int * self.nI)()
        m = (C.c_uint8 * self.N)()
        self.lib.pg_info_set(self.ctx, I, m)
        self.I = np.array(I[:], dtype=np.int32)
        self.inI = np.array(m[:], dtype=np.uint8)

    def close(self):
        if self.ctx:
            self.lib.pg_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, self.lib.pg_last_error(self.ctx).decode()))

    def decode_llr(self, llr, packed=False):
        """llr (B,N) float32, float64 or float16 host array -> (u_hat (B,N) uint8 | (B,N/32) uint32, flags (B,) uint32)."""
        llr = np.ascontiguousarray(llr)
        if llr.dtype not in (np.float32, np.float64, np.float16):
            llr = llr.astype(np.float64)
        llr = llr.reshape(-1, self.N)
        B = llr.shape[0]
        fmt = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.float16): 2}[llr.dtype]   # PG_LLR_F32 / F64 / F16
        flags = np.zeros(B, dtype=np.uint32)
        if packed:
            out = np.zeros((B, self.N // 32), dtype=np.uint32)
            rc = self.lib.pg_decode_llr_packed(self.ctx, llr.ctypes.data, fmt, B, out.ctypes.data, flags.ctypes.data)
        else:
            out = np.zeros((B, self.N), dtype=np.uint8)
            rc = self.lib.pg_decode_llr(self.ctx, llr.ctypes.data, fmt, B, out.ctypes.data, flags.ctypes.data)
        self._check(rc, "pg_decode_llr")
        return out, flags

    def decode_llr_counted(self, llr, u_true, acc=None):
        """pg_decode_llr_counted: -> (u_hat (B,N) uint8, frame_err (B,) uint16, counters)"""
        llr = np.ascontiguousarray(llr)
        if llr.dtype not in (np.float32, np.float64):
            llr = llr.astype(np.float64)
        llr = llr.reshape(-1, self.N)
        B = llr.shape[0]
        u_true = np.ascontiguousarray(u_true, dtype=np.uint8).reshape(B, self.N)
        out = np.zeros((B, self.N), dtype=np.uint8)
        fe = np.zeros(B, dtype=np.uint16)
        acc = acc if acc is not None else PgCounters()
        rc = self.lib.pg_decode_llr_counted(self.ctx, llr.ctypes.data, int(llr.dtype == np.float64), B, u_true.ctypes.data, out.ctypes.data, C.byref(acc), fe.ctypes.data)
        self._check(rc, "pg_decode_llr_counted")
        return out, fe, acc

    def channel(self, ebn0_db, first_frame, B):
        """-> (llr (B,N) float32|float64, u (B,N) uint8) exactly as pg_simulate would feed the decoder."""
        llr = np.zeros((B, self.N), dtype=np.float64 if self.f64 else np.float32)
        u = np.zeros((B, self.N), dtype=np.uint8)
        self._check(self.lib.pg_channel(self.ctx, float(ebn0_db), int(first_frame), B, llr.ctypes.data, u.ctypes.data), "pg_channel")
        return llr, u

    def simulate_batch(self, ebn0_db, first_frame, B, want_frame_err=False, acc=None):
        acc = acc if acc is not None else PgCounters()
        fe = np.zeros(B, dtype=np.uint16) if want_frame_err else None
        rc = self.lib.pg_simulate_batch(self.ctx, float(ebn0_db), int(first_frame), B, C.byref(acc), fe.ctypes.data if want_frame_err else None)
        self._check(rc, "pg_simulate_batch")
        return acc, fe

    def simulate(self, ebn0_db, first_frame=0, target_err_blocks=0, max_frames=0, exact_stop=True):
        out = PgCounters()
        rc = self.lib.pg_simulate(self.ctx, float(ebn0_db), int(first_frame), int(target_err_blocks), int(max_frames), int(bool(exact_stop)), C.byref(out))
        self._check(rc, "pg_simulate")
        return out

    def simulate_stats(self):
        """(rounds launched, all-reduces issued) by the last simulate() call"""
        r, a = C.c_uint64(), C.c_uint64()
        self._check(self.lib.pg_simulate_stats(self.ctx, C.byref(r), C.byref(a)), "pg_simulate_stats")
        return int(r.value), int(a.value)

    # ---- device-pointer calls (pointers are plain integers, e.g. torch.Tensor.data_ptr())
    def channel_device(self, ebn0_db, first_frame, B, d_llr, d_u_packed):
        self._check(self.lib.pg_channel_device(self.ctx, float(ebn0_db), int(first_frame), B, d_llr, d_u_packed), "pg_channel_device")

    def decode_count_device(self, d_llr, B, d_truth, d_u_hat=None, d_info=None):
        self._check(self.lib.pg_decode_count_device(self.ctx, d_llr, int(self.f64), B, d_truth, d_u_hat, d_info), "pg_decode_count_device")

    def decode_llr_device(self, d_llr, llr_is_f64, B, d_u_hat=None, d_flags=None):
        self._check(self.lib.pg_decode_llr_device(self.ctx, d_llr, int(llr_is_f64), B, d_u_hat, d_flags), "pg_decode_llr_device")

    def counters_read(self, reset=False):
        out = PgCounters()
        self._check(self.lib.pg_counters_read(self.ctx, C.byref(out), int(reset)), "pg_counters_read")
        return out

    def decode_llr_host_ptr(self, llr_ptr, llr_is_f64, B, out_packed_ptr, flags_ptr=None):
        """pg_decode_llr_packed on raw host pointers (pinned buffers owned by the caller); llr_is_f64: False/True or a PG_LLR_* code"""
        self._check(self.lib.pg_decode_llr_packed(self.ctx, llr_ptr, int(llr_is_f64), B, out_packed_ptr, flags_ptr), "pg_decode_llr_packed")

    def stream_ptr(self):
        return int(self.lib.pg_stream(self.ctx) or 0)

    def allreduce_counters(self, c):
        self._check(self.lib.pg_allreduce_counters(self.ctx, C.byref(c)), "pg_allreduce_counters")
        return c

    def bpr_config(self, samples):
        arr = (C.c_int * len(samples))(*samples)
        self._check(self.lib.pg_bpr_config(self.ctx, arr, len(samples)), "pg_bpr_config")
        self._bpr_ns = len(samples)

    def bpr_read(self):
        n = int(np.log2(self.N))
        E = (C.c_uint64 * (self._bpr_ns * (n + 1)))()
        self._check(self.lib.pg_bpr_read(self.ctx, E), "pg_bpr_read")
        return np.array(E[:], dtype=np.uint64).reshape(self._bpr_ns, n + 1)

    def comm_init(self, id128: bytes):
        buf = C.create_string_buffer(id128, 128)
        self._check(self.lib.pg_comm_init(self.ctx, buf), "pg_comm_init")

    def wave_frames(self):
        return int(self.lib.pg_wave_frames(self.ctx))

    def sync(self):
        self._check(self.lib.pg_sync(self.ctx), "pg_sync")

    def launches(self):
        return int(self.lib.pg_kernel_launches(self.ctx))

    def last_kernel_ms(self):
        d, c = C.c_float(), C.c_float()
        self._check(self.lib.pg_last_kernel_ms(self.ctx, C.byref(d), C.byref(c)), "pg_last_kernel_ms")
        return d.value, c.value


def comm_unique_id():
    buf = C.create_string_buffer(128)
    rc = load_library().pg_comm_unique_id(buf)
    if rc != 0:
        raise RuntimeError("pg_comm_unique_id failed: %s" % load_library().pg_last_error(None).decode())
    return buf.raw
