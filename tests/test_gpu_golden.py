"""GPU: the CUDA path against the committed golden vectors of the COMPILED REFERENCE (tests/golden/*.npz), the channel
kernel against the reference's frame construction, and the Monte-Carlo loop's stopping rule."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PROGS = ["SC_128", "SC_1024", "SC_128_fag", "SCL_128", "SCL_128_fag", "CASCL_128", "SCL_1024", "CASCL_1024_L8",
         "CASCL_1024_sys", "BP_128", "BP_128_fag", "BP_1024"]


def load(prog):
    z = np.load(os.path.join(GOLD, prog + ".npz"))
    N = int(z["N"])
    return z, z["llr"], np.unpackbits(z["u"], axis=1, bitorder="little")[:, :N], np.unpackbits(z["u_hat"], axis=1, bitorder="little")[:, :N]


@pytest.mark.parametrize("prog", PROGS)
def test_fp64_kernel_reproduces_reference_decisions(prog):
    from polardecoding_b200 import Engine
    z, llr, u, uh = load(prog)
    eng = Engine(prog, real="f64")
    got, flags = eng.decode_llr(llr)            # float32 LLRs, converted on the device
    assert (got == uh).all(), "%d frames differ from the compiled reference" % int((got != uh).any(1).sum())
    packed, _ = eng.decode_llr(llr.astype(np.float64), packed=True)
    assert (np.unpackbits(packed.view(np.uint8), axis=1, bitorder="little") == uh).all()
    eng.close()


@pytest.mark.parametrize("prog", ["SC_128", "CASCL_128", "CASCL_1024_L8", "CASCL_1024_sys", "BP_1024"])
def test_channel_kernel_frames(prog):
    """payload/CRC/placement are the reference's (PN phase m = frame*(K%63) mod 63), noise has the right moments,
    and a frame does not depend on batch split or on where the batch starts"""
    from polardecoding_b200 import Engine
    o = Oracle(prog)
    for real in ("f32", "f64"):
        eng = Engine(prog, real=real, seed=77)
        first, B = 1000, 300
        llr, u = eng.channel(2.0, first, B)
        pn = np.zeros(63, dtype=np.int32)
        o.lib.po_pn63(pn.ctypes.data_as(C.POINTER(C.c_int)))
        want = np.zeros(o.N, dtype=np.int32)
        for f in (0, 1, 62, 63, 299):
            m = ((first + f) * (o.K % 63)) % 63
            o.lib.po_make_u(C.byref(o.code), pn.ctypes.data_as(C.POINTER(C.c_int)), m, want.ctypes.data_as(C.POINTER(C.c_int)))
            assert (u[f] == want).all(), (prog, f)
        x = u.astype(np.int32).copy()
        s = 1
        while s < o.N:
            xr = x.reshape(B, -1, 2, s)
            xr[:, :, 0, :] ^= xr[:, :, 1, :]
            s *= 2
        sigma = 10 ** (2.0 / -20)
        z = (llr.astype(np.float64) * sigma * sigma / 2 - (1 - 2 * x)) / sigma      # recovered unit noise
        assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01 and abs((z ** 3).mean()) < 0.03 and abs((z ** 4).mean() - 3) < 0.1
        llr2, u2 = eng.channel(2.0, first + 100, 37)
        assert (llr2 == llr[100:137]).all() and (u2 == u[100:137]).all()
        eng.close()


@pytest.mark.parametrize("prog,ebn0", [("SC_128", 2.0), ("CASCL_128", 1.5), ("BP_128", 2.0), ("CASCL_1024_L8", 1.0)])
def test_simulation_counts_and_exact_stop(prog, ebn0):
    from polardecoding_b200 import Engine
    B = 6000 if "128" in prog else 1500
    eng = Engine(prog, real="f64", seed=5)
    acc, fe = eng.simulate_batch(ebn0, 0, B, want_frame_err=True)
    # the fused path equals channel -> decode -> compare done by hand
    llr, u = eng.channel(ebn0, 0, B)
    got, _ = eng.decode_llr(llr)
    cnt_from = eng.params.count_from
    pos = eng.I[cnt_from:]
    nerr = (got[:, pos] != u[:, pos]).sum(1)
    assert (fe == nerr).all()
    assert acc.frames == B and acc.err_blocks == int((nerr > 0).sum()) and acc.err_bits == int(nerr.sum())
    assert acc.err_blocks >= 30, "test needs errors to be meaningful"
    # reference stopping rule: run = index of the target-th erroneous frame + 1, independent of batching
    for target in (1, 7, 25):
        r = eng.simulate(ebn0, first_frame=0, target_err_blocks=target, exact_stop=True)
        idx = np.nonzero(nerr > 0)[0][target - 1]
        assert (r.frames, r.err_blocks, r.err_bits) == (idx + 1, target, int(nerr[: idx + 1].sum()))
    r = eng.simulate(ebn0, first_frame=0, max_frames=1234)
    assert (r.frames, r.err_blocks) == (1234, int((nerr[:1234] > 0).sum()))
    eng.close()


def test_fer_matches_reference_tables():
    """BLER of the published captures within the 95 % interval of both samples"""
    from polardecoding_b200 import Engine
    for prog, ebn0, ref_bler, ref_err, frames in [("SC_128", 2.0, 0.1414, 100, 40000), ("CASCL_128", 2.0, 0.04182, 200, 60000),
                                                    ("CASCL_1024_L8", 1.5, 0.07130, 200, 30000), ("BP_1024", 2.0, 0.03292, 200, 16000),
                                                    ("SCL_1024", 1.5, 0.04873, 50, 30000)]:
        eng = Engine(prog, real="f32", seed=2024, data_mode=1)
        r = eng.simulate(ebn0, max_frames=frames)
        bler = r.err_blocks / r.frames
        sd = ref_bler * np.sqrt(1.0 / ref_err + 1.0 / max(1, r.err_blocks))
        print("%s %.1f dB: BLER %.5f (%d errors / %d frames) reference %.5f" % (prog, ebn0, bler, r.err_blocks, r.frames, ref_bler))
        assert abs(bler - ref_bler) < 2.6 * sd, (prog, bler, ref_bler)
        eng.close()


def test_bpr_statistic_matches_reference():
    """BPr_128.c:418-568: per-stage hard decision + re-encode error counts E[sample][stage], against the compiled
    reference's E (golden) and against the oracle on fresh frames; also with the fixed-point stop (same statistic)."""
    from polardecoding_b200 import Engine
    z, llr, u, uh = load("BPr_128")
    samples = [int(v) for v in z["samples"]]
    for early in (0, 1):
        eng = Engine("BPr_128", real="f64", bp_early_stop=early)
        eng.bpr_config(samples)
        got, fe, acc = eng.decode_llr_counted(llr, u)
        assert (got == uh).all()
        E = eng.bpr_read()
        assert (E == z["E"]).all(), (early, E, z["E"])
        assert acc.frames == len(llr) and (fe == (got != u)[:, eng.I].sum(1)).all()
        eng.close()
    o = Oracle("BPr_128")
    uu, l2 = o.frames_ref_stream(2.0, 40, seed=99)
    _, Eo = o.bpr(l2, uu, samples)
    eng = Engine("BPr_128", real="f64")
    eng.bpr_config(samples)
    eng.decode_llr_counted(l2, uu)
    assert (eng.bpr_read() == Eo).all()
    eng.close()
